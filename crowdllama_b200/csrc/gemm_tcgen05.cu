// gemm_tcgen05.cu — prefill / batched-decode weight projections on the 5th-gen tensor cores.
//
//   Y[t][n] (+)= sum_k X[t][k] * W[n][k]        X bf16 [T][K], W bf16 [N][K] (both K-major), Y fp32
//
// This is the "weight projections as tcgen05/TMA tensor-core tiles for prefill" row of SURVEY.md
// §8a (kernel table); upstream counterpart is ggml-cuda's cuBLAS / mma.sync mmq prefill path.
//
// Mapping (swap-AB so small token counts waste N, not M):
//   UMMA A operand = W tile  [128 rows n ][64 k]  -> accumulator lanes  (TMEM lane  = n)
//   UMMA B operand = X tile  [BT  rows t ][64 k]  -> accumulator columns (TMEM column = t)
//   accumulator D[128][BT] fp32 in TMEM, double buffered (2*BT columns).
//   MT = 2 (prefill, T >= 256): one CTA owns a 256 (n) x 256 (t) tile as two M = 128 MMAs per k-step that share the token
//   operand.  The 128 x 256 tile needs 96 B/clk/SM of operands from L2 at full MMA rate = 14.2 KB/clk for the chip, more
//   than twice what the L2 slices deliver (~6.3 KB/clk, B300_MICROARCH.md "LTS throughput cap"); the 256 x 256 tile needs
//   64 B/clk/SM.  Its two accumulators fill TMEM (512 columns), so the epilogue no longer overlaps the next tile — with
//   K >= 4096 a tile is >= 65K clk of MMA against ~2K clk of epilogue.
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0   TMA producer  : cp.async.bulk.tensor.2d (128B swizzle) into an NST-stage mbarrier ring
//   warp 1   MMA issuer    : tcgen05.mma.cta_group::1.kind::f16, tcgen05.commit -> frees smem stage /
//                            signals the epilogue; also owns tcgen05.alloc / dealloc
//   warps 2-5 epilogue     : tcgen05.ld 32x32b.x32 -> registers -> coalesced fp32 stores
//                            (lane = n, so a warp writes 32 consecutive n of one token: 128 B)
// Tile order: consecutive tiles share the W tile (t fastest), so W streams from HBM once and X
// (<= 32 MB) stays L2-resident.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace cl {

namespace {

using namespace tc;

struct GemmParams {
  float* Y;
  int T, N, K, ldy;
  int accumulate_into_y;  // 1: Y += result (residual add in place)
  int k_splits;           // > 1: split-K; split s writes its partial to Y + s * T * ldy (the consumer sums in fixed order)
  int pre_stages;         // PDL launches only: W tiles of the first stages requested before griddepcontrol.wait (0 = none)
  GemmEpi epi;            // fused prefill epilogues (EPI template parameter != 0): see kernels.h
};
enum { EPI_F32 = 0, EPI_SILU = 1, EPI_ROPE = 2 };

template <int BT, int NST, int MT, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int BMT = BM * MT;                      // W rows per tile
  constexpr uint32_t A_BYTES = BMT * BK * 2;        // 16 KB per M = 128 sub-tile
  constexpr uint32_t B_BYTES = BT * BK * 2;
  constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  constexpr int NACC = (2 * MT * BT <= 512) ? 2 : 1;   // accumulator sets in TMEM
  constexpr uint32_t TMEM_COLS = NACC * MT * BT < 32 ? 32 : NACC * MT * BT;
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns: power of two <= 512");
  // dynamic smem base is only guaranteed 16-byte aligned: round up to 1024 for the 128B swizzle
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)NST * STAGE);
  uint64_t* empty = full + NST;
  uint64_t* tmem_full = empty + NST;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_t = (p.T + BT - 1) / BT, num_n = (p.N + BMT - 1) / BMT;
  const int ksp = p.k_splits > 1 ? p.k_splits : 1;
  const int num_tiles = num_t * num_n * ksp;          // split index is the slowest dimension
  const int num_kb_all = (p.K + BK - 1) / BK;
  const int kb_per = (num_kb_all + ksp - 1) / ksp;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_x);
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }   // [NACC] used
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // no-ops unless launched with the programmatic-serialization attribute (batched decode step)

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // X is the previous kernel's output; W is constant: the W halves of the first p.pre_stages ring stages are
      // requested before griddepcontrol.wait (while the glue kernel in front still runs), their X halves after it.
      int pre = 0;
      if (p.pre_stages > 0 && (int)blockIdx.x < num_tiles) {
        const int tile = blockIdx.x;
        const int bt_ = tile % (num_t * num_n), ks = tile / (num_t * num_n);
        const int n0 = (bt_ / num_t) * BMT;
        const int kb0 = ks * kb_per, kb1 = min(num_kb_all, kb0 + kb_per);
        pre = min(min(p.pre_stages, NST), kb1 - kb0);
        for (int i = 0; i < pre; ++i) {            // fresh barriers: every stage is empty
          mbar_arrive_expect_tx(&full[i], STAGE);
          tma_load_2d(base + (size_t)i * STAGE, &map_w, (kb0 + i) * BK, n0, &full[i]);
        }
      }
      pdl_wait();
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int bt_ = tile % (num_t * num_n), ks = tile / (num_t * num_n);
        const int n0 = (bt_ / num_t) * BMT, t0 = (bt_ % num_t) * BT;
        const int kb0 = ks * kb_per, kb1 = min(num_kb_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          uint8_t* sa = base + (size_t)stage * STAGE;
          if (pre > 0) {                           // W of this stage is already on its way
            --pre;
            tma_load_2d(sa + A_BYTES, &map_x, kb * BK, t0, &full[stage]);
          } else {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], STAGE);
            tma_load_2d(sa, &map_w, kb * BK, n0, &full[stage]);
            tma_load_2d(sa + A_BYTES, &map_x, kb * BK, t0, &full[stage]);
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(BM, BT);
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * MT * BT);
      const int ks = tile / (num_t * num_n);
      const int kb0 = ks * kb_per, num_kb = max(0, min(num_kb_all, kb0 + kb_per) - kb0);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(base + (size_t)stage * STAGE);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)  // advance 32 bytes (= 2 x 16 B units) per UMMA_K inside the swizzle atom
#pragma unroll
            for (int m = 0; m < MT; ++m)     // M sub-tiles: 128 W rows = 16 KB (1024 x 16 B units) apart, same token operand
              umma_f16(d_addr + (uint32_t)(m * BT), adesc + (uint64_t)(2 * k + m * 1024), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quarter = warp % 4 =================
    pdl_wait();              // Y / the split-K workspace may still be read by the previous kernel
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int bt_ = tile % (num_t * num_n), ks = tile / (num_t * num_n);
      const int n0 = (bt_ / num_t) * BMT, t0 = (bt_ % num_t) * BT;
      float* Yb = p.Y + (size_t)ks * p.T * p.ldy;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int n = n0 + m * BM + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BT; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + m) * BT + c0), v);
          tmem_ld_wait();
          if constexpr (EPI == EPI_F32) {
            if (n < p.N) {
              if (p.accumulate_into_y) {
                // residual add in place: ALL 32 loads first, then the stores.  (A load-add-store per token serialises
                // on the load latency — the compiler cannot reorder the next load above a store it cannot prove
                // disjoint — and made the o / down projections epilogue-bound: 564 us instead of 114 us per launch.)
                float y[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  const int t = t0 + c0 + c;
                  y[c] = t < p.T ? __ldcg(Yb + (size_t)t * p.ldy + n) : 0.f;
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  const int t = t0 + c0 + c;
                  if (t < p.T) Yb[(size_t)t * p.ldy + n] = y[c] + __uint_as_float(v[c]);
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  const int t = t0 + c0 + c;
                  if (t < p.T) Yb[(size_t)t * p.ldy + n] = __uint_as_float(v[c]);
                }
              }
            }
          } else if constexpr (EPI == EPI_SILU) {
            // rows are interleaved (2i = gate_i, 2i+1 = up_i): the lane pair (2j, 2j+1) holds one SwiGLU input pair per
            // token.  Two tokens per iteration so that EVERY lane does useful work: the even lane finishes token c (it
            // owns g, receives u), the odd lane token c + 1 (owns u, receives g) — one shuffle, one SiLU, one store per
            // lane and pair of tokens; per token the 16 lanes of one parity write 16 consecutive bf16 = a full 32-byte
            // sector.  (One token per iteration with the odd lanes idle made this epilogue longer than the tile's MMAs.)
            const int i = n >> 1, odd = lane & 1;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float a0 = __uint_as_float(v[c]), a1 = __uint_as_float(v[c + 1]);
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);   // odd sends u[c], even sends g[c + 1]
              const float gt = odd ? recv : a0, up = odd ? a1 : recv;
              const int t = t0 + c0 + c + odd;
              if (n < p.N && t < p.T)
                p.epi.act[(size_t)t * p.epi.ld_act + i] = __float2bfloat16_rn(__fdividef(gt, 1.0f + __expf(-gt)) * up);
            }
          } else {
            // q|k|v rows are rope-pair-interleaved per head (row 2j = dim j, row 2j+1 = dim j + 64; BM = head_dim = 128, so
            // a 128-row sub-tile is exactly one head): the lane pair holds one rotation pair per token.  Rotate at the
            // token's position, round to bf16, write q to q_out[t][head][dim] and k / v straight into the paged cache.
            const int head = (n0 + m * BM) >> 7, w = n & 127, j = w >> 1;
            const int dim = (lane & 1) ? j + 64 : j;
            const GemmEpi& e = p.epi;
            const bool rot = head < e.n_heads + e.n_kv, isq = head < e.n_heads;
            // table entries of the 32 tokens first (independent loads in flight), then rotate / round / store
            float2 cs[32];
            int pg[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int t = t0 + c0 + c, pos = e.pos0 + (t < p.T ? t : 0);
              cs[c] = rot ? __ldg(e.rope + (size_t)pos * 64 + j) : make_float2(1.f, 0.f);
              pg[c] = isq ? 0 : __ldg(e.block_table + pos / e.page_size);
            }
            const int g = rot ? head - e.n_heads : head - e.n_heads - e.n_kv;
            __nv_bfloat16* pool = rot ? e.kpool : e.vpool;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float mine = __uint_as_float(v[c]);
              const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
              const int t = t0 + c0 + c;
              // even lane: v0 = mine (dim j), v1 = other -> v0 cos - v1 sin; odd lane: v1 = mine (dim j + 64) -> v1 cos + v0 sin
              const float val = (lane & 1) ? mine * cs[c].x + other * cs[c].y : mine * cs[c].x - other * cs[c].y;
              if (t < p.T && n < p.N) {
                if (isq) {
                  e.q_out[(size_t)t * e.q_dim + head * 128 + dim] = __float2bfloat16_rn(val);
                } else {
                  const int off = (e.pos0 + t) % e.page_size;
                  pool[(((size_t)pg[c] * e.n_kv + g) * e.page_size + off) * 128 + dim] = __float2bfloat16_rn(val);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows][K] tensor, box = {64 k, box_rows}, 128B swizzle, OOB -> zeros
bool make_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// 2-D bf16 tensor map for KV pools etc.: [rows][cols] row-major, box {box_cols, box_rows}, 128B swizzle
bool make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

template <int BT, int NST, int MT = 1, int EPI = EPI_F32>
cudaError_t launch_inst(const CUtensorMap& mw, const CUtensorMap& mx, const GemmParams& p, cudaStream_t st, bool pdl) {
  auto kern = gemm_tcgen05_kernel<BT, NST, MT, EPI>;
  constexpr size_t smem = (size_t)NST * (MT * BM * BK * 2 + BT * BK * 2) + 1024 + 256;
  static PerDeviceOnce attr;
  if (attr.pending()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr.mark();
  }
  const int tiles = ((p.T + BT - 1) / BT) * ((p.N + BM * MT - 1) / (BM * MT)) * (p.k_splits > 1 ? p.k_splits : 1);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, mw, mx, p);
}

}  // namespace

bool gemm_tcgen05_supported(int T, int N, int K) { return T > 0 && N > 0 && K > 0 && K % 8 == 0 && get_encode() != nullptr; }

int launch_gemm_bf16(const __nv_bfloat16* X, const __nv_bfloat16* W, float* Y, const float* resid, int T, int N, int K,
                     cudaStream_t st, int k_splits, bool pdl) {
  if (!gemm_tcgen05_supported(T, N, K)) return -1;
  if (resid && resid != Y) return -1;  // residual add is in place
  const int BT = T > 128 ? 256 : T > 64 ? 128 : T > 32 ? 64 : 32;
  // 256 x 256 tiles (two M sub-tiles per CTA, CL_GEMM_MT=2) cut the operand stream from L2 by a third, but measured
  // SLOWER on the Llama-3-8B shapes at T = 4096 (r2d: 1122 vs 1216 TFLOP/s sustained per layer; only the K = 14336
  // down-projection gains, 1316 vs 1132): wave quantisation (256 instead of 512 tiles on 148 SMs) and the lost
  // epilogue overlap cost more than the L2 traffic saves.  Kept opt-in for the measurements.
  static const int mt_env = getenv("CL_GEMM_MT") ? atoi(getenv("CL_GEMM_MT")) : 1;
  // 128-row token tile (batched decode steps of 65-128 sequences): the X tile is as large as a 128-row W tile, so the
  // operand stream from L2 is twice the weight stream; two M sub-tiles per CTA would share one X tile (wide projections
  // only: with N / 256 tiles the narrow ones no longer fill the machine).  Measured (r2ab, ctx 256): B = 80 / 100 / 128 =
  // 4.96 / 5.24 / 5.93 ms against 4.90 / 5.18 / 5.86 with one sub-tile — the L2 operand stream is not what bounds these
  // steps.  Opt-in (CL_GEMM_MT_B128=1), parity-checked by test_llama3_8b_layers_wide_batch under that switch.
  static const int mt128_env = getenv("CL_GEMM_MT_B128") ? atoi(getenv("CL_GEMM_MT_B128")) : 0;
  const int MT = (BT == 256 && k_splits <= 1 && N >= 2 * BM && mt_env >= 2) ? 2
               : (BT == 128 && mt128_env && N >= 16384 && N % (2 * BM) == 0) ? 2 : 1;
  CUtensorMap mw, mx;
  if (!make_map(&mw, W, N, K, BM * MT) || !make_map(&mx, X, T, K, BT)) return -1;
  if (k_splits > 1 && (resid || (K + BK - 1) / BK < k_splits)) return -1;   // every split needs >= 1 k-block (an empty split would never signal its epilogue)
  if (k_splits > 1 && ((K + BK - 1) / BK + k_splits - 1) / k_splits * (k_splits - 1) >= (K + BK - 1) / BK) return -1;
  // measured (r2p / r2q, batched decode step at ctx 1024): B = 8 3.819 ms with no early W, 3.758 / 3.734 / 3.719 with 4 / 6 / 8
  // stages; B = 32 4.525 -> 4.487.  The whole ring it is (the round-1 loss came from the glue kernels of that time).
  static const int pre_env = getenv("CL_GEMM_PRE") ? atoi(getenv("CL_GEMM_PRE")) : 8;
  GemmParams p{Y, T, N, K, N, resid ? 1 : 0, k_splits, pdl ? pre_env : 0, GemmEpi()};
  cudaError_t e;
  switch (BT) {
    case 256: e = MT == 2 ? launch_inst<256, 3, 2>(mw, mx, p, st, pdl) : launch_inst<256, 4>(mw, mx, p, st, pdl); break;
    case 128: e = MT == 2 ? launch_inst<128, 4, 2>(mw, mx, p, st, pdl) : launch_inst<128, 6>(mw, mx, p, st, pdl); break;
    case 64: e = launch_inst<64, 8>(mw, mx, p, st, pdl); break;
    default: e = launch_inst<32, 8>(mw, mx, p, st, pdl); break;
  }
  return e == cudaSuccess ? 1 : -1;
}

// Prefill projections with a fused epilogue (no fp32 round trip through HBM): kind 1 = gate|up -> bf16(SiLU(g) * u),
// kind 2 = q|k|v -> RoPE + bf16 + q buffer / paged KV cache.  One CTA per 128 x BT tile (no split-K, no M pairing).
int launch_gemm_bf16_epi(const __nv_bfloat16* X, const __nv_bfloat16* W, int T, int N, int K, const GemmEpi& epi, cudaStream_t st) {
  if (!gemm_tcgen05_supported(T, N, K) || (epi.kind != 1 && epi.kind != 2) || (N & 1)) return -1;
  if (epi.kind == 2 && (N % 128 || N != (epi.n_heads + 2 * epi.n_kv) * 128)) return -1;   // head_dim 128 = one M sub-tile per head
  const int BT = T > 128 ? 256 : T > 64 ? 128 : T > 32 ? 64 : 32;
  CUtensorMap mw, mx;
  if (!make_map(&mw, W, N, K, BM) || !make_map(&mx, X, T, K, BT)) return -1;
  GemmParams p{nullptr, T, N, K, N, 0, 1, 0, epi};
  cudaError_t e;
  if (epi.kind == 1) {
    switch (BT) {
      case 256: e = launch_inst<256, 4, 1, EPI_SILU>(mw, mx, p, st, false); break;
      case 128: e = launch_inst<128, 6, 1, EPI_SILU>(mw, mx, p, st, false); break;
      case 64: e = launch_inst<64, 8, 1, EPI_SILU>(mw, mx, p, st, false); break;
      default: e = launch_inst<32, 8, 1, EPI_SILU>(mw, mx, p, st, false); break;
    }
  } else {
    switch (BT) {
      case 256: e = launch_inst<256, 4, 1, EPI_ROPE>(mw, mx, p, st, false); break;
      case 128: e = launch_inst<128, 6, 1, EPI_ROPE>(mw, mx, p, st, false); break;
      case 64: e = launch_inst<64, 8, 1, EPI_ROPE>(mw, mx, p, st, false); break;
      default: e = launch_inst<32, 8, 1, EPI_ROPE>(mw, mx, p, st, false); break;
    }
  }
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace cl
