// gemm_skinny.cu — weight projections of the BATCHED decode step (2..32 sequences advance together).
//
//   Y[t][n] = sum_k X[t][k] * W[n][k]       X bf16 [T <= 32][K], W bf16 [N][K], fp32 accumulation
//
// With a handful of token columns the projection is a pure weight stream (HBM-bound), so the kernel is built
// like the single-token GEMV ring — but the arithmetic runs on the tensor cores (tcgen05, swap-AB: the 128
// weight rows of a tile are the UMMA M dimension, the tokens the N = 32 dimension; accumulators in TMEM).
//
//   * work unit = (128-row weight tile, k-split): `kb_per` k-blocks of 64 = a 64..128 KB slice of W.  Units are
//     claimed DYNAMICALLY from a global counter (SMs do not get equal HBM bandwidth; 1.5 waves of static tiles
//     left the old split-K launch at 4.4 TB/s).  The producer lane claims, the MMA and epilogue warps follow
//     through a small shared-memory queue.
//   * split-K partials go to an L2-resident workspace; the LAST split to arrive at a (weight tile, 32-row quarter)
//     counter sums all partials in fixed split order — deterministic, whoever is last — and runs the fused
//     epilogue.  Each epilogue warp owns one quarter end to end (no CTA barrier) and consumes the arrival
//     atomic's result one unit later, so the round trip never stalls the unit pipeline.  Counters re-arm
//     themselves (graph replay).
//   * fused epilogues (same contracts as the GEMV epilogues, kernels.h): plain store / residual add into the
//     fp32 stream / SiLU(gate)*up -> bf16 / RoPE + bf16 + q out + paged KV append.
//   * PDL: weights never depend on the previous kernel, so the producer fills the ring before
//     griddepcontrol.wait; only the X tiles (and every epilogue access) wait.
//
// Reference counterpart: ggml-cuda's mmvq / small-batch mul_mat path behind ollama's runner (not vendored;
// call site /root/reference/pkg/crowdllama/api.go:108-160).
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace cl {

namespace {

using namespace tc;

constexpr int BT = 32;                      // token columns (UMMA N)
constexpr int NST = 10;                     // ring stages
constexpr uint32_t A_BYTES = BM * BK * 2;   // 16 KB of W per stage
constexpr uint32_t B_BYTES = BT * BK * 2;   // 4 KB of X per stage (rows >= T are zero-filled by TMA, no traffic)
constexpr uint32_t STAGE = A_BYTES + B_BYTES;
constexpr int UQ = 16;                      // unit queue depth
constexpr int NACC = 8;                     // TMEM accumulators (32 columns each): the MMA warp may run 8 units ahead of the epilogue
constexpr uint32_t TMEM_COLS = NACC * BT;

struct SkinnyParams {
  int T, N, K, ksp, kb_per, num_kb, num_n, epi, tile_major;
  float* part;            // [ksp][T][N] split-K partials (ksp > 1)
  unsigned* sched;        // [0] unit counter, [1] CTAs done (self-cleaning)
  unsigned* tile_cnt;     // [num_n][4] split arrivals per (weight tile, 32-row quarter) (self-cleaning)
  float* y; int ldy;      // SK_STORE / SK_RESID target rows: y[row(t)][n], row(t) = slots ? slots[t] : t
  const int* slots;
  __nv_bfloat16* act;     // SK_GATEUP: act[t][n / 2]
  QkvEpi qkv;             // SK_QKV
  float* q_out; int q_stride;
  long long* dbg;         // diagnostics: globaltimer stamps of CTA 0 (nullptr = off)
};

#define SK_STAMP(k) do { if (p.dbg && blockIdx.x == 0) p.dbg[k] = gtime_ns(); } while (0)

__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }

// per-token metadata of the step, staged in shared memory once per CTA (the epilogue must not chase
// slots[] -> pos[] -> block_table[] pointers per element: those dependent loads serialise behind the stores)
struct TokMeta {
  int slot[BT];
  int pos[BT];
  long long kvrow[BT];    // SK_QKV: (page * n_kv) * page_size + off  ->  row base of kv head 0 in the pool, in tokens
};

// final values v[t] of row n for all tokens -> fused epilogue.  Called by all 32 lanes of an epilogue warp (lane = row).
// Loads first (independent, one round trip), stores after.
template <int EPI>
__device__ __forceinline__ void skinny_epilogue(const SkinnyParams& p, const TokMeta& m, int n, const float (&v)[BT], int lane) {
  const bool ok = n < p.N;
  if (EPI == SK_STORE) {
#pragma unroll
    for (int t = 0; t < BT; ++t)
      if (t < p.T && ok) p.y[(size_t)m.slot[t] * p.ldy + n] = v[t];
  } else if (EPI == SK_RESID) {
    float hv[BT];
#pragma unroll
    for (int t = 0; t < BT; ++t)
      if (t < p.T && ok) hv[t] = __ldcg(p.y + (size_t)m.slot[t] * p.ldy + n);
#pragma unroll
    for (int t = 0; t < BT; ++t)
      if (t < p.T && ok) p.y[(size_t)m.slot[t] * p.ldy + n] = hv[t] + v[t];   // single writer per element
  } else if (EPI == SK_GATEUP) {
#pragma unroll
    for (int t = 0; t < BT; ++t) {
      if (t < p.T) {
        const float up = __shfl_down_sync(0xffffffffu, v[t], 1);      // rows (2i, 2i+1) = (gate_i, up_i)
        if (!(lane & 1) && ok) p.act[(size_t)t * (p.N >> 1) + (n >> 1)] = __float2bfloat16_rn(v[t] / (1.0f + __expf(-v[t])) * up);
      }
    }
  } else {   // SK_QKV: rows (2j, 2j+1) of a head = RoPE pair (dim j, dim j + head_dim/2)
    const QkvEpi& e = p.qkv;
    const int HD = e.head_dim, half = HD >> 1;
    const int pr = n >> 1, hh = pr / half, j = pr - hh * half;
    const bool rot = hh < e.n_heads + e.n_kv;
    float2 cs[BT];
#pragma unroll
    for (int t = 0; t < BT; ++t)
      if (t < p.T && ok && rot && !(lane & 1)) cs[t] = __ldg(e.rope + (size_t)m.pos[t] * half + j);
#pragma unroll
    for (int t = 0; t < BT; ++t) {
      if (t < p.T) {
        const float v0 = v[t], v1 = __shfl_down_sync(0xffffffffu, v[t], 1);
        if (!(lane & 1) && ok) {
          if (rot) {
            const float r0 = bf16_round(v0 * cs[t].x - v1 * cs[t].y), r1 = bf16_round(v1 * cs[t].x + v0 * cs[t].y);
            if (hh < e.n_heads) {
              float* q = p.q_out + (size_t)m.slot[t] * p.q_stride + (size_t)hh * HD;
              q[j] = r0; q[j + half] = r1;
            } else {
              const size_t base = ((size_t)m.kvrow[t] + (size_t)(hh - e.n_heads) * e.page_size) * HD;
              e.kpool[base + j] = __float2bfloat16_rn(r0);
              e.kpool[base + j + half] = __float2bfloat16_rn(r1);
            }
          } else {
            const size_t base = ((size_t)m.kvrow[t] + (size_t)(hh - e.n_heads - e.n_kv) * e.page_size) * HD;
            e.vpool[base + j] = __float2bfloat16_rn(v0);
            e.vpool[base + j + half] = __float2bfloat16_rn(v1);
          }
        }
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const SkinnyParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)NST * STAGE);
  uint64_t* empty = full + NST;
  uint64_t* tmem_full = empty + NST;         // [NACC]
  uint64_t* tmem_empty = tmem_full + NACC;   // [NACC]
  uint64_t* uq_full = tmem_empty + NACC;     // [UQ]
  uint64_t* uq_empty = uq_full + UQ;      // [UQ]
  int* uq = reinterpret_cast<int*>(uq_empty + UQ);   // [UQ]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(uq + UQ);
  TokMeta& meta = *reinterpret_cast<TokMeta*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);
  int* wait_list = reinterpret_cast<int*>(&meta + 1);   // [4 epilogue warps][16]

  const long long t_entry = p.dbg ? gtime_ns() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = p.num_n * p.ksp;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_x);
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    for (int i = 0; i < UQ; ++i) { mbar_init(&uq_full[i], 1); mbar_init(&uq_empty[i], 5); }   // consumers: MMA warp + 4 epilogue warps
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
  pdl_launch_dependents();
  if (threadIdx.x == 0) SK_STAMP(0);
  if (p.dbg && threadIdx.x == 0) { p.dbg[16 + 2 * blockIdx.x] = t_entry; p.dbg[16 + 2 * blockIdx.x + 1] = gtime_ns(); }

  if (warp == 0) {
    // ================= producer: claims units, streams W (+ X) tiles =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int qi = 0; uint32_t qphase = 0;
      bool waited = false;
      int pend[NST]; int n_pend = 0;          // stages whose X tile is still owed (issued before griddepcontrol.wait)
      // unit claims: the first claim takes enough consecutive units to fill the ring at once; afterwards one unit
      // per claim, issued just before the LAST k-block of the current unit (the atomic's round trip hides behind
      // that copy, and a claimed unit starts at once — claiming a whole unit ahead made the slowest CTA finish up
      // to two units late).
      int local = 0;
      auto claim = [&](int cnt) -> int {
        if (p.sched) return (int)atomicAdd(p.sched, (unsigned)cnt);
        const int u = ((int)blockIdx.x + local * (int)gridDim.x) * cnt;   // static order (cnt is constant then)
        ++local;
        return u;
      };
      const int c0 = p.sched ? max(1, min(4, (NST + p.kb_per - 1) / p.kb_per)) : 1;
      int u = claim(c0), u_end = u + c0;     // [u, u_end): claimed, not yet issued
      SK_STAMP(1);
      while (true) {
        const bool valid = u < num_units;
        mbar_wait(&uq_empty[qi], qphase ^ 1u);
        uq[qi] = valid ? u : -1;
        mbar_arrive(&uq_full[qi]);
        if (++qi == UQ) { qi = 0; qphase ^= 1u; }
        if (!valid) break;
        const int tile = p.tile_major ? u / p.ksp : u % p.num_n, ks = p.tile_major ? u % p.ksp : u / p.num_n;
        const int n0 = tile * BM;
        const int kb0 = ks * p.kb_per, kb1 = min(p.num_kb, kb0 + p.kb_per);
        int nxt = u + 1;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kb == kb1 - 1 && nxt == u_end) { nxt = claim(1); u_end = nxt + 1; }
          if (!waited && n_pend == NST) {       // ring full of W tiles: now the X tiles need the previous kernel
            SK_STAMP(2);
            pdl_wait();
            SK_STAMP(3);
            waited = true;
            for (int i = 0; i < n_pend; ++i) {
              const int st = pend[i] & 0xff, kbx = pend[i] >> 8;
              tma_load_2d(base + (size_t)st * STAGE + A_BYTES, &map_x, kbx * BK, 0, &full[st]);
            }
            n_pend = 0;
          }
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], STAGE);
          uint8_t* sa = base + (size_t)stage * STAGE;
          tma_load_2d(sa, &map_w, kb * BK, n0, &full[stage]);
          if (waited) tma_load_2d(sa + A_BYTES, &map_x, kb * BK, 0, &full[stage]);
          else pend[n_pend++] = stage | (kb << 8);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        u = nxt;
      }
      SK_STAMP(10);
      if (!waited) {
        pdl_wait();
        for (int i = 0; i < n_pend; ++i) {
          const int st = pend[i] & 0xff, kbx = pend[i] >> 8;
          tma_load_2d(base + (size_t)st * STAGE + A_BYTES, &map_x, kbx * BK, 0, &full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(BM, BT);
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    int qi = 0; uint32_t qphase = 0;
    bool mma_seen = false, unit_seen = false;
    while (true) {
      mbar_wait(&uq_full[qi], qphase);
      const int u = uq[qi];
      __syncwarp();
      if (lane == 0) mbar_arrive(&uq_empty[qi]);
      if (++qi == UQ) { qi = 0; qphase ^= 1u; }
      if (u < 0) break;
      const int ks = p.tile_major ? u % p.ksp : u / p.num_n;
      const int kb0 = ks * p.kb_per, num_kb = min(p.num_kb, kb0 + p.kb_per) - kb0;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BT);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0 && !mma_seen) { SK_STAMP(4); mma_seen = true; }
        if (lane == 0) {
          const uint32_t sa = smem_u32(base + (size_t)stage * STAGE);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_f16(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) { umma_commit(&tmem_full[acc]); if (!unit_seen) { SK_STAMP(5); unit_seen = true; } }
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lane quarter = warp % 4 =================
    // Every warp runs its own 32-row quarter of the tile end to end — no CTA-level barrier.  Split-K: the warp
    // publishes its partial rows, then lane 0 arrives at the (tile, quarter) counter with an acq_rel atomic whose
    // result is only consumed one unit LATER (the round trip hides behind the next unit's TMEM load and stores).
    // The warp that finds itself last sums all splits in fixed order and runs the fused epilogue.
    pdl_wait();   // every global access below may touch buffers the previous kernel still reads or writes
    const int q = warp & 3;
    if (warp == 2 && lane < p.T) {
      const int slot = p.slots ? p.slots[lane] : lane;
      meta.slot[lane] = slot;
      if (EPI == SK_QKV) {
        const QkvEpi& e = p.qkv;
        const int pos = e.pos[slot];
        const int page = e.block_tables[(size_t)slot * e.bt_stride + pos / e.page_size];
        meta.pos[lane] = pos;
        meta.kvrow[lane] = ((long long)page * e.n_kv) * e.page_size + pos % e.page_size;
      }
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    int acc = 0; uint32_t acc_phase = 0;
    int qi = 0; uint32_t qphase = 0;
    // Split-K protocol.  Every unit stores its partial rows and ARRIVES at the (tile, quarter) counter; arrivals are
    // batched — one release (membar.gpu costs the warp ~1 us whether or not anything is outstanding) publishes up
    // to AB units, lanes 0..AB-1 each bump one counter.  The reduction of a tile is owned by the warp that processed
    // its LAST split index (not by whoever arrives last: the slowest SM would then be last everywhere, inherit every
    // reduction and become slower still).  The owner keeps the tile on a small wait list, polls the counter after
    // each batch (all entries in one round trip) and reduces when it reads ksp; at the end it drains the list.
    constexpr int AB = 4, WL = 16;
    int arr_tile[AB], n_arr = 0;
    int* wl = wait_list + (warp - 2) * WL;   // shared memory, private to this warp
    int n_wait = 0;
    bool epi_seen = false;
    auto reduce_tile = [&](int tile) {
      if (lane == 0) p.tile_cnt[tile * 4 + q] = 0u;   // every split has arrived: re-arm for the next launch (graph replay)
      __syncwarp();                                   // the polling lane's acquire ordered before the other lanes' loads
      const int n = tile * BM + q * 32 + lane;
      // fixed-order sum over the splits, 8 splits x 8 tokens = 64 independent L2 loads in flight per lane
      float sum[BT];
#pragma unroll
      for (int t = 0; t < BT; ++t) sum[t] = 0.f;
      const bool ok = n < p.N;
      const size_t sstride = (size_t)p.T * p.N;
#pragma unroll
      for (int t0 = 0; t0 < BT; t0 += 8) {
        if (t0 < p.T) {
          for (int s0 = 0; s0 < p.ksp; s0 += 8) {
            float tmp[8][8];
#pragma unroll
            for (int ss = 0; ss < 8; ++ss)
#pragma unroll
              for (int tt = 0; tt < 8; ++tt)
                tmp[ss][tt] = (ok && s0 + ss < p.ksp && t0 + tt < p.T) ? ldcg_f(p.part + (size_t)(s0 + ss) * sstride + (size_t)(t0 + tt) * p.N + n) : 0.f;
#pragma unroll
            for (int ss = 0; ss < 8; ++ss)
#pragma unroll
              for (int tt = 0; tt < 8; ++tt) sum[t0 + tt] += tmp[ss][tt];
          }
        }
      }
      skinny_epilogue<EPI>(p, meta, n, sum, lane);
    };
    auto arrive = [&]() {          // publish the partial rows of the units stored since the last arrive
      if (n_arr == 0) return;
      __syncwarp();                // all lanes' partial stores ordered before the releasing lanes' reds
      int my_tile = -1;
#pragma unroll
      for (int j = 0; j < AB; ++j) if (j < n_arr && lane == j) my_tile = arr_tile[j];
      if (my_tile >= 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.tile_cnt + my_tile * 4 + q) : "memory");
      n_arr = 0;
    };
    auto poll = [&](bool drain) {  // reduce every owned tile whose splits have all arrived
      long long t0 = 0;
      while (n_wait > 0) {
        __syncwarp();
        const int mine = lane < n_wait ? wl[lane] : -1;
        const unsigned cnt = mine >= 0 ? ld_acquire_u32(p.tile_cnt + mine * 4 + q) : 0u;
        const unsigned ready = __ballot_sync(0xffffffffu, mine >= 0 && cnt == (unsigned)p.ksp);
        if (ready) {
          int keep = 0;
          for (int i = 0; i < n_wait; ++i) {     // warp-uniform
            const int tile = wl[i];
            if ((ready >> i) & 1u) reduce_tile(tile);
            else { __syncwarp(); if (lane == 0) wl[keep] = tile; ++keep; }
          }
          n_wait = keep;
          __syncwarp();
        }
        if (!drain) break;
        if (n_wait > 0 && !ready) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > (1ll << 31)) __trap();
        }
      }
    };
    while (true) {
      mbar_wait(&uq_full[qi], qphase);
      const int u = uq[qi];
      __syncwarp();
      if (lane == 0) mbar_arrive(&uq_empty[qi]);
      if (++qi == UQ) { qi = 0; qphase ^= 1u; }
      if (u < 0) break;
      const int tile = p.tile_major ? u / p.ksp : u % p.num_n, ks = p.tile_major ? u % p.ksp : u / p.num_n;
      const int n = tile * BM + q * 32 + lane;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 && lane == 0 && !epi_seen) { SK_STAMP(6); epi_seen = true; }
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BT), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);   // the accumulator is in registers: hand TMEM back early
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      if (p.ksp == 1) {
        float vf[BT];
#pragma unroll
        for (int t = 0; t < BT; ++t) vf[t] = __uint_as_float(v[t]);
        skinny_epilogue<EPI>(p, meta, n, vf, lane);
        continue;
      }
      // store this unit's partial rows; every AB units: consume the previous batch of arrivals, publish this one
      if (n_arr == AB) { arrive(); poll(false); }
      if (n < p.N) {
        float* pp = p.part + (size_t)ks * p.T * p.N + n;
#pragma unroll
        for (int t = 0; t < BT; ++t)
          if (t < p.T) __stcg(pp + (size_t)t * p.N, __uint_as_float(v[t]));
      }
#pragma unroll
      for (int j = 0; j < AB; ++j) if (j == n_arr) arr_tile[j] = tile;
      ++n_arr;
      if (ks == p.ksp - 1) {       // this warp owns the tile's reduction
        if (n_wait == WL) { arrive(); poll(true); }
        if (lane == 0) wl[n_wait] = tile;
        ++n_wait;
      }
    }
    if (warp == 2 && lane == 0) SK_STAMP(7);
    arrive();
    poll(true);
    if (warp == 2 && lane == 0) SK_STAMP(8);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SK_STAMP(9);
  if (p.dbg && threadIdx.x == 0) p.dbg[16 + 2 * 160 + blockIdx.x] = gtime_ns();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if (p.sched && threadIdx.x == 0) {
    const unsigned done = atomicAdd(p.sched + 1, 1u);
    if (done == gridDim.x - 1) { p.sched[0] = 0u; p.sched[1] = 0u; }   // every CTA has stopped claiming: re-arm
  }
}

template <typename Kern, typename Args>
cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, bool pdl, const CUtensorMap& mw, const CUtensorMap& mx,
                       const Args& args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, mw, mx, args);
}

template <int EPI>
cudaError_t launch_skinny_inst(const CUtensorMap& mw, const CUtensorMap& mx, const SkinnyParams& p, cudaStream_t st, bool pdl) {
  auto kern = gemm_skinny_kernel<EPI>;
  constexpr size_t smem = (size_t)NST * STAGE + 1024 + (2 * NST + 2 * NACC + 2 * UQ) * 8 + UQ * 4 + 64 + sizeof(TokMeta) + 4 * 16 * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int units = p.num_n * p.ksp;
  const int grid = units < sm_count() ? units : sm_count();
  return launch_pdl(kern, grid, 192, smem, st, pdl, mw, mx, p);
}

}  // namespace

// k-splits of a skinny projection: units of `target_kb` k-blocks (64 columns each), every split non-empty
int skinny_splits(int K, int target_kb) {
  const int nkb = (K + BK - 1) / BK;
  int kb_per = target_kb < 1 ? 1 : target_kb;
  if (kb_per > nkb) kb_per = nkb;
  return (nkb + kb_per - 1) / kb_per;
}
size_t skinny_counter_words(int N) { return 2 + 4 * (size_t)((N + BM - 1) / BM); }

int launch_gemm_skinny(const SkinnyArgs& a, cudaStream_t st, bool pdl) {
  if (a.T < 1 || a.T > BT || a.K % 8 != 0 || a.N < 1) return -1;
  if (a.epi != SK_STORE && (a.N % 2)) return -1;
  CUtensorMap mw, mx;
  if (!make_tmap_2d_bf16(&mw, a.W, (uint64_t)a.N, (uint64_t)a.K, BK, BM) || !make_tmap_2d_bf16(&mx, a.X, (uint64_t)a.T, (uint64_t)a.K, BK, BT))
    return -1;
  SkinnyParams p{};
  p.T = a.T; p.N = a.N; p.K = a.K;
  p.num_kb = (a.K + BK - 1) / BK;
  p.ksp = a.k_splits < 1 ? 1 : a.k_splits;
  p.kb_per = (p.num_kb + p.ksp - 1) / p.ksp;
  p.ksp = (p.num_kb + p.kb_per - 1) / p.kb_per;          // no empty split
  p.num_n = (a.N + BM - 1) / BM;
  p.epi = a.epi;
  { static const int tm = getenv("CL_SKINNY_TILE_MAJOR") ? atoi(getenv("CL_SKINNY_TILE_MAJOR")) : 0; p.tile_major = tm; }
  p.part = a.part;
  { static const int dyn = getenv("CL_SKINNY_STATIC") ? !atoi(getenv("CL_SKINNY_STATIC")) : 1; p.sched = dyn ? a.counters : nullptr; }
  p.tile_cnt = a.counters ? a.counters + 2 : nullptr;
  if (p.ksp > 1 && (!p.part || !p.tile_cnt)) return -1;
  p.y = a.y; p.ldy = a.ldy; p.slots = a.slots; p.act = a.act; p.qkv = a.qkv; p.q_out = a.q_out; p.q_stride = a.q_stride; p.dbg = a.dbg;
  cudaError_t e;
  switch (a.epi) {
    case SK_STORE: e = launch_skinny_inst<SK_STORE>(mw, mx, p, st, pdl); break;
    case SK_RESID: e = launch_skinny_inst<SK_RESID>(mw, mx, p, st, pdl); break;
    case SK_GATEUP: e = launch_skinny_inst<SK_GATEUP>(mw, mx, p, st, pdl); break;
    case SK_QKV: e = launch_skinny_inst<SK_QKV>(mw, mx, p, st, pdl); break;
    default: return -1;
  }
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace cl
