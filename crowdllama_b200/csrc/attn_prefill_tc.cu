// attn_prefill_tc.cu — causal GQA prefill attention over the paged KV cache on the 5th-gen tensor cores.
//
// "prefill causal attention | T x T per head, d = 128 | tensor pipe" row of SURVEY.md §8a; upstream counterpart:
// ggml-cuda's flash_attn_ext mma tile kernel behind the Ollama server (/root/reference/pkg/crowdllama/api.go:129-139).
// Replaces the mma.sync kernel of prefill_kernels.cu for head_dim 128 / page 32 (that one stays for head_dim 64).
//
// One work item = one query head x 256 consecutive query rows = two 128-row tiles (A, B) that share every K/V block.
// Persistent CTAs (one per SM) take items round-robin from a heaviest-first order.  Per 128-token KV block j and tile X:
//     S_X = Q_X K_j^T        tcgen05.mma  M 128 (query rows = TMEM lanes) x N 128 (tokens) x K 128 (head dim), fp32 in TMEM
//     P_X = exp2(S_X - m)    softmax warps: tcgen05.ld one row per thread, online max / sum in registers,
//                            P rounded to bf16 into shared memory in the K-major 128B-swizzled UMMA layout
//     O_X += P_X V_j         tcgen05.mma  M 128 x N 128 (head dim) x K 128 (tokens); V straight from the paged cache
//                            as an MN-major (token rows, dims contiguous) 128B-swizzled operand; O stays in TMEM
// TMEM: S_A | S_B | O_A | O_B = 4 x 128 fp32 columns (all 512).  The running maximum is only raised when a block exceeds
// it by more than 2^8 (then the owning warp rescales its 32 rows of O in TMEM with tcgen05.ld / tcgen05.st): softmax is
// invariant to the reference point, p <= 256 is harmless in bf16 / fp32, and after the first blocks of a row the
// correction almost never runs.
// Warp roles (320 threads): warp 0 = TMA producer (Q tiles; K / V pages through a 3 x 32 KB ring, 2-D tensor maps over the
// whole pool, 128B swizzle), warp 1 = MMA issuer (one elected thread; tcgen05.commit -> mbarriers), warps 2-5 = softmax
// group of tile A, warps 6-9 = tile B.  While one group runs its exponentials the tensor core works for the other.
// Numerics (cl-llama v1): q, K, V bf16; scores, softmax, accumulation fp32; P rounded to bf16 before P.V
// (the documented deviation of the prefill path); output rounded to bf16 = the o-projection's GEMM operand.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace cl {

namespace {

using namespace tc;

constexpr int HD = 128;                 // head dim
constexpr int TQ = 128;                 // query rows per tile
constexpr int TKV = 128;                // tokens per KV block
constexpr int PAGE = 32;
constexpr uint32_t TILE_BYTES = 32768;  // a [128][128] bf16 operand tile = two [128 rows][128 B] swizzled slabs
constexpr uint32_t SLAB = 16384;
constexpr int NRING = 3;
constexpr int MAX_BT = 320;             // block-table entries staged in shared memory (10240 tokens)
constexpr float RESCALE_LOG2 = 8.0f;

struct AttnTcParams {
  const int* block_table;               // this sequence
  int n_bt;                             // entries of block_table that are valid (pages covering pos0 + T tokens)
  int pos0, T, n_heads, n_kv;
  long long layer_row0;                 // first row of this layer in the pool-wide K / V tensor maps
  __nv_bfloat16* out;                   // [T][n_heads * HD]
  float scale2;                         // 1/sqrt(HD) * log2(e)
};

// MN-major, 128B-swizzled operand (V: rows = tokens = K of the MMA, 64 dims = one 128-byte row):
// SBO = 1024 B between 8-token groups, LBO = 16384 B between the two 64-dim halves
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(SLAB >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC_S = make_idesc(TQ, TKV);                  // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = make_idesc(TQ, HD) | (1u << 16);     // A = P (K-major), B = V (MN-major)

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// work item -> (first query row, head); heaviest (latest rows) first, heads fastest
struct Item { int q0, head, nblk[2]; };
__device__ __forceinline__ bool get_item(const AttnTcParams& p, int idx, int n_pairs, Item* it) {
  if (idx >= n_pairs * p.n_heads) return false;
  const int pr = n_pairs - 1 - idx / p.n_heads;
  it->head = idx % p.n_heads;
  it->q0 = pr * 2 * TQ;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int r0 = it->q0 + x * TQ;
    const int last = min(r0 + TQ, p.T) - 1;                         // last real query row of the tile
    it->nblk[x] = r0 < p.T ? (p.pos0 + last) / TKV + 1 : 0;         // KV blocks 0 .. nblk-1 hold keys <= its position
  }
  return true;
}

__global__ void __launch_bounds__(320, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = base;                                  // [2][32 KB]
  uint8_t* p_s = base + 2 * TILE_BYTES;                 // [2][32 KB]
  uint8_t* ring = base + 4 * TILE_BYTES;                // [NRING][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (4 + NRING) * TILE_BYTES);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* s_full = bars + 4;        // [2]
  uint64_t* p_full = bars + 6;        // [2]
  uint64_t* o_full = bars + 8;        // [2]
  uint64_t* o_empty = bars + 10;      // [2]
  uint64_t* kv_full = bars + 12;      // [NRING]
  uint64_t* kv_empty = bars + 12 + NRING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * NRING);
  int* bt_s = reinterpret_cast<int*>(tmem_slot + 2);    // [MAX_BT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (p.T + 2 * TQ - 1) / (2 * TQ);
  const int g_rep = p.n_heads / p.n_kv;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v);
    for (int x = 0; x < 2; ++x) {
      mbar_init(&q_full[x], 1); mbar_init(&q_empty[x], 1); mbar_init(&s_full[x], 1);
      mbar_init(&p_full[x], 128); mbar_init(&o_full[x], 1); mbar_init(&o_empty[x], 128);
    }
    for (int i = 0; i < NRING; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_bt && i < MAX_BT; i += blockDim.x) bt_s[i] = p.block_table[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      uint32_t ring_it = 0, na[2] = {0u, 0u};                       // na[x]: items so far in which tile x was active
      Item it;
      for (int idx = blockIdx.x; get_item(p, idx, n_pairs, &it); idx += gridDim.x) {
        const int g = it.head / g_rep;
        for (int x = 0; x < 2; ++x) {
          if (it.nblk[x] == 0) continue;
          mbar_wait(&q_empty[x], (na[x] & 1u) ^ 1u);                // the previous active item's last S of this tile has completed
          ++na[x];
          mbar_arrive_expect_tx(&q_full[x], TILE_BYTES);
          uint8_t* dst = q_s + x * TILE_BYTES;
          tma_load_2d(dst, &map_q, it.head * HD, it.q0 + x * TQ, &q_full[x]);
          tma_load_2d(dst + SLAB, &map_q, it.head * HD + 64, it.q0 + x * TQ, &q_full[x]);
        }
        const int nmax = it.nblk[1] > it.nblk[0] ? it.nblk[1] : it.nblk[0];
        for (int j = 0; j < nmax; ++j) {
          for (int kv = 0; kv < 2; ++kv) {                          // ring order: K_0, V_0, K_1, V_1, ...
            const int slot = (int)(ring_it % NRING);
            mbar_wait(&kv_empty[slot], ((ring_it / NRING) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&kv_full[slot], TILE_BYTES);
            uint8_t* dst = ring + (size_t)slot * TILE_BYTES;
            const CUtensorMap* mp = kv ? &map_v : &map_k;
#pragma unroll
            for (int i = 0; i < TKV / PAGE; ++i) {
              int pi = j * (TKV / PAGE) + i;
              if (pi >= p.n_bt) pi = 0;                             // past the sequence: any valid page (its keys are masked, P = 0)
              const long long row = p.layer_row0 + ((long long)bt_s[pi] * p.n_kv + g) * PAGE;
              tma_load_2d(dst + i * 4096, mp, 0, (int)row, &kv_full[slot]);
              tma_load_2d(dst + SLAB + i * 4096, mp, 64, (int)row, &kv_full[slot]);
            }
            ++ring_it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =======================================
    uint32_t ring_it = 0, np[2] = {0u, 0u}, na[2] = {0u, 0u};       // np[x]: p_full phases consumed; na[x]: active items so far
    Item it;
    for (int idx = blockIdx.x; get_item(p, idx, n_pairs, &it); idx += gridDim.x) {
      const int nmax = it.nblk[1] > it.nblk[0] ? it.nblk[1] : it.nblk[0];
      // S of block jj for both tiles from ring entry K_jj (ring index r_k); releases the K slot and, after a tile's
      // last S, its Q tile
      auto issue_s = [&](int jj, uint32_t r_k, int only) {
        const uint32_t kaddr = smem_u32(ring + (size_t)(r_k % NRING) * TILE_BYTES);
        for (int x = 0; x < 2; ++x) {
          if (only >= 0 && x != only) continue;
          if (jj >= it.nblk[x]) continue;
          if (lane == 0) {
            const uint32_t qaddr = smem_u32(q_s + x * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < HD / UK; ++k) {
              const uint32_t off = (uint32_t)(k >> 2) * SLAB + (uint32_t)(k & 3) * 32u;
              umma_f16(tmem_base + (uint32_t)(x * TKV), make_smem_desc(qaddr + off), make_smem_desc(kaddr + off), IDESC_S, k ? 1u : 0u);
            }
            umma_commit(&s_full[x]);
            if (jj == it.nblk[x] - 1) umma_commit(&q_empty[x]);
          }
          __syncwarp();
        }
      };
      uint32_t oe_par[2];
      for (int x = 0; x < 2; ++x) {
        oe_par[x] = (na[x] & 1u) ^ 1u;
        if (it.nblk[x]) { mbar_wait(&q_full[x], na[x] & 1u); ++na[x]; }
      }
      tc_fence_after();
      // prologue: S(0) of both tiles
      uint32_t r_k = ring_it;                                       // ring index of K_0
      mbar_wait(&kv_full[r_k % NRING], (r_k / NRING) & 1u);
      tc_fence_after();
      issue_s(0, r_k, -1);
      if (lane == 0) umma_commit(&kv_empty[r_k % NRING]);
      __syncwarp();
      for (int j = 0; j < nmax; ++j) {
        const uint32_t r_v = ring_it + 2 * j + 1, r_kn = r_v + 1;   // V_j, K_{j+1}
        mbar_wait(&kv_full[r_v % NRING], (r_v / NRING) & 1u);
        if (j + 1 < nmax) mbar_wait(&kv_full[r_kn % NRING], (r_kn / NRING) & 1u);
        tc_fence_after();
        const uint32_t vaddr = smem_u32(ring + (size_t)(r_v % NRING) * TILE_BYTES);
        for (int x = 0; x < 2; ++x) {
          if (j >= it.nblk[x]) continue;
          mbar_wait(&p_full[x], np[x] & 1u);                        // P_j written, S_x read, O_x(j-1) settled
          ++np[x];
          if (j == 0) mbar_wait(&o_empty[x], oe_par[x]);            // the previous active item's epilogue has read O_x
          tc_fence_after();
          issue_s(j + 1, r_kn, x);                                  // first: the other group's next scores start early
          if (lane == 0) {
            const uint32_t paddr = smem_u32(p_s + x * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < TKV / UK; ++k) {
              const uint64_t adesc = make_smem_desc(paddr + (uint32_t)(k >> 2) * SLAB + (uint32_t)(k & 3) * 32u);
              const uint64_t bdesc = make_smem_desc_mn(vaddr + (uint32_t)k * (UK * 128u));
              umma_f16(tmem_base + (uint32_t)(2 * TKV + x * HD), adesc, bdesc, IDESC_PV, (j | k) ? 1u : 0u);
            }
            umma_commit(&o_full[x]);
          }
          __syncwarp();
        }
        if (lane == 0) {
          umma_commit(&kv_empty[r_v % NRING]);
          if (j + 1 < nmax) umma_commit(&kv_empty[r_kn % NRING]);
        }
        __syncwarp();
      }
      ring_it += 2u * (uint32_t)nmax;
    }
  } else {
    // ======================================= softmax groups =======================================
    const int x = (warp - 2) >> 2;                                  // tile A: warps 2-5, tile B: warps 6-9
    const int qd = warp & 3;                                        // TMEM lane quarter this warp may access
    const int r = qd * 32 + lane;                                   // row of the tile = TMEM lane
    const uint32_t t_s = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(x * TKV);
    const uint32_t t_o = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(2 * TKV + x * HD);
    uint8_t* prow = p_s + x * TILE_BYTES + r * 128;
    const int sw = r & 7;
    uint32_t ns = 0, no = 0;                                        // s_full / o_full phases consumed
    Item it;
    for (int idx = blockIdx.x; get_item(p, idx, n_pairs, &it); idx += gridDim.x) {
      const int nb = it.nblk[x];
      if (nb == 0) continue;
      const int qrow = it.q0 + x * TQ + r;
      const int qabs = p.pos0 + qrow;
      float m_ref = 0.f, l = 0.f;
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&s_full[x], ns & 1u);
        ++ns;
        tc_fence_after();
        const int k0 = j * TKV;
        const bool masked = k0 + TKV - 1 > p.pos0 + it.q0 + x * TQ;   // some key of the block lies above some row's diagonal
        float mb = -INFINITY, sum = 0.f;
        // half a row of S (64 tokens): scaled scores, running block maximum, p = 2^(s - mref) rounded to bf16 pairs
        auto compute_half = [&](int hh, float mref, uint32_t (&pk)[32]) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = hh * 2 + cc;
            uint32_t v[32];
            tmem_ld32(t_s + (uint32_t)(c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a = __uint_as_float(v[i]) * p.scale2, b = __uint_as_float(v[i + 1]) * p.scale2;
              if (masked) {
                a = (k0 + c * 32 + i <= qabs) ? a : -INFINITY;
                b = (k0 + c * 32 + i + 1 <= qabs) ? b : -INFINITY;
              }
              mb = fmaxf(mb, fmaxf(a, b));
              const float p0 = ex2_approx(a - mref), p1 = ex2_approx(b - mref);
              sum += p0 + p1;
              pk[(cc * 32 + i) >> 1] = pack_bf16(p0, p1);
            }
          }
        };
        // -> shared memory, K-major 128B-swizzled UMMA layout: token c of the block lives in slab c/64 (= hh),
        // 16-byte chunk ((c%64)/8) ^ (row & 7) of this row's 128 bytes
        auto store_half = [&](int hh, const uint32_t (&pk)[32]) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(prow + hh * SLAB + ((ch ^ sw) << 4)) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        };
        float corr = 1.0f;
        uint32_t pk[32];
        if (j == 0) {
          // no reference yet: maximum first (block 0 holds key 0 <= every position: finite).  The P buffer is free: this
          // group waited for the previous item's last P.V in its epilogue.
          float m0 = -INFINITY;
#pragma unroll
          for (int c = 0; c < TKV / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(t_s + (uint32_t)(c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float a = __uint_as_float(v[i]) * p.scale2;
              if (!masked || k0 + c * 32 + i <= qabs) m0 = fmaxf(m0, a);
            }
          }
          m_ref = m0;
          compute_half(0, m_ref, pk); store_half(0, pk);
          compute_half(1, m_ref, pk); store_half(1, pk);
        } else {
          // optimistic single pass against the running reference: the first half's exponentials overlap P.V of block
          // j-1; its stores need the P buffer, i.e. P.V(j-1) complete
          compute_half(0, m_ref, pk);
          mbar_wait(&o_full[x], no & 1u);                           // P.V of block j-1 done: P buffer free, O_x stable
          ++no;
          tc_fence_after();
          store_half(0, pk);
          compute_half(1, m_ref, pk); store_half(1, pk);
          const bool raise = mb > m_ref + RESCALE_LOG2;
          if (__any_sync(0xffffffffu, raise)) {                     // rare: a row outgrew its reference by 2^8
            if (raise) { corr = ex2_approx(m_ref - mb); m_ref = mb; }
            mb = -INFINITY; sum = 0.f;
            compute_half(0, m_ref, pk); store_half(0, pk);          // redo the block against the new reference
            compute_half(1, m_ref, pk); store_half(1, pk);
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {                     // and move this warp's 32 rows of O to it
              uint32_t v[32];
              tmem_ld32(t_o + (uint32_t)(c * 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
              tmem_st32(t_o + (uint32_t)(c * 32), v);
            }
            tmem_st_wait();
          }
        }
        l = l * corr + sum;
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
      }
      // epilogue: O / l -> bf16 -> out[qrow][head * HD ..]
      mbar_wait(&o_full[x], no & 1u);
      ++no;
      tc_fence_after();
      const float inv = 1.0f / l;
      __nv_bfloat16* orow = p.out + (size_t)qrow * ((size_t)p.n_heads * HD) + (size_t)it.head * HD;
#pragma unroll 1
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_o + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (qrow < p.T) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + i) = w;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_empty[x]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

constexpr size_t kSmem = (size_t)(4 + NRING) * TILE_BYTES + (12 + 2 * NRING) * 8 + 8 + MAX_BT * 4 + 1024;

}  // namespace

bool attn_prefill_tc_supported(int n_heads, int n_kv, int head_dim, int page_size, int pos0, int T) {
  const char* ev = getenv("CL_PREFILL_ATTN_TC");     // 0: keep the mma.sync kernel (prefill_kernels.cu)
  const bool off = ev && atoi(ev) == 0;
  return !off && head_dim == HD && page_size == PAGE && n_kv > 0 && n_heads % n_kv == 0 && T > 0 &&
         (pos0 + T + PAGE - 1) / PAGE <= MAX_BT;
}

// q: [T][n_heads*128] bf16 (roped); K / V through the pool-wide tensor maps (box {64, 32}, 128B swizzle);
// layer_row0 = first row of this layer in those maps
int launch_attn_prefill_tc(const AttnPrefillArgs& a, const CUtensorMap& kmap, const CUtensorMap& vmap, long long layer_row0, cudaStream_t st) {
  if (!attn_prefill_tc_supported(a.n_heads, a.n_kv, a.head_dim, a.page_size, a.pos0, a.T)) return -1;
  CUtensorMap qmap;
  if (!make_tmap_2d_bf16(&qmap, a.q, (uint64_t)a.T, (uint64_t)a.n_heads * HD, 64, TQ)) return -1;
  static PerDeviceOnce attr;
  if (attr.pending()) {
    if (cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return -1;
    attr.mark();
  }
  AttnTcParams p;
  p.block_table = a.block_table; p.n_bt = (a.pos0 + a.T + PAGE - 1) / PAGE; p.pos0 = a.pos0; p.T = a.T; p.n_heads = a.n_heads; p.n_kv = a.n_kv;
  p.layer_row0 = layer_row0; p.out = a.out; p.scale2 = 1.4426950408889634f / sqrtf((float)HD);
  const int items = ((a.T + 2 * TQ - 1) / (2 * TQ)) * a.n_heads;
  const int grid = items < sm_count() ? items : sm_count();
  attn_prefill_tc_kernel<<<grid, 320, kSmem, st>>>(qmap, kmap, vmap, p);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace cl
