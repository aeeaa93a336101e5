// common.cuh — device helpers shared by the sm_100a kernels of libclengine.so.
// Numerics contract "cl-llama v1": see DESIGN.md §3 (bf16 weights / KV, fp32 residual stream,
// bf16-rounded GEMV inputs, fp32 accumulation).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cl {

constexpr int kWarp = 32;

// ---- bf16 <-> fp32 (bit tricks; RNE rounding identical to oracle/llama_oracle.c) -------------
__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf16_round(float f) {
  return __uint_as_float((uint32_t)f32_to_bf16_bits(f) << 16);
}

// ---- streaming 128-bit global load (weights / KV are read once: keep them out of L1) ---------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// 256-bit streaming load (sm_100: LDG.E.NA.EFL2.256) — 32 bytes per lane, L2 evict-first
struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg_stream256(const void* p) {
  u32x8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
// dot of 16 bf16 (one 256-bit load) with 16 fp32 held in shared memory (4 x float4)
__device__ __forceinline__ float dot16(const u32x8& w, const float4* x4, float acc) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 xv = x4[q];
    acc = fmaf(bf16_lo(w.v[2 * q]), xv.x, acc); acc = fmaf(bf16_hi(w.v[2 * q]), xv.y, acc);
    acc = fmaf(bf16_lo(w.v[2 * q + 1]), xv.z, acc); acc = fmaf(bf16_hi(w.v[2 * q + 1]), xv.w, acc);
  }
  return acc;
}
__device__ __forceinline__ uint4 ldg_stream_keep(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// dot of 8 bf16 (one uint4) with 8 fp32, fp32 FMA chain
__device__ __forceinline__ float dot8(const uint4& w, const float4& xa, const float4& xb, float acc) {
  acc = fmaf(bf16_lo(w.x), xa.x, acc); acc = fmaf(bf16_hi(w.x), xa.y, acc);
  acc = fmaf(bf16_lo(w.y), xa.z, acc); acc = fmaf(bf16_hi(w.y), xa.w, acc);
  acc = fmaf(bf16_lo(w.z), xb.x, acc); acc = fmaf(bf16_hi(w.z), xb.y, acc);
  acc = fmaf(bf16_lo(w.w), xb.z, acc); acc = fmaf(bf16_hi(w.w), xb.w, acc);
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- flag-based dependencies between the kernels of one token step -----------------------------
// griddepcontrol.wait releases a dependent grid only after the primary grid has fully completed and
// flushed (~2-3 us).  Inside the token step every kernel instead publishes "my outputs are visible"
// by a release-increment of a per-node counter, and its consumer (already resident thanks to the early
// PDL trigger) acquires it by polling: the dependent starts ~0.5 us after the last producer CTA.
// The counters are zeroed by step_bump_kernel at the end of every step (full stream dependency).
typedef unsigned long long u64;
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// lane 0 of the calling warp polls; the warp continues when the counter has reached `expect`
__device__ __forceinline__ void wait_counter_warp(const unsigned* cnt, unsigned expect) {
  if ((threadIdx.x & 31) == 0) {
    const long long t0 = clock64();
    while (ld_acquire_u32(cnt) < expect) {
      __nanosleep(40);
      if (clock64() - t0 > (1ll << 31)) __trap();   // ~1 s: protocol bug -> kernel error, never a hung box
    }
  }
  __syncwarp();
}
// call after a CTA-wide barrier that follows the last global write of the CTA
__device__ __forceinline__ void signal_counter(unsigned* cnt) {
  __threadfence();
  atomicAdd(cnt, 1u);
}
__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- mbarrier + bulk async copy (TMA 1-D) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) { __trap(); }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// 2-D TMA tile load (cp.async.bulk.tensor -> UTMALDG); c0 = innermost coordinate
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// warp-level tensor-core helpers (mma.sync m16n8k16 bf16 -> fp32, ldmatrix)
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- counter-based synthetic weights (must match oracle/llama_oracle.c oc_synth_int) ----------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ int synth_int(uint64_t seed, int key, uint64_t index) {
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + (((uint64_t)(uint32_t)key << 40) | index);
  uint32_t r = (uint32_t)mix64(x);
  return (int)((r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + (r >> 24)) - 510;
}

}  // namespace cl
