// tc_common.cuh -- PTX wrappers shared by the warp-specialised tcgen05 kernels (lstm_tc.cu): mbarriers, TMA tensor loads,
// tcgen05 MMA / commit / TMEM loads, proxy fences.  sm_100a only.  (gemm_tc.cu predates this header and keeps its own
// copies of the few wrappers it needs.)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cb200 {
namespace tc {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// a lost arrival must surface as an error (trap -> launch failure), never as a hang: waits are bounded in wall time
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;
struct SpinGuard {
  unsigned spin = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void tick() {
    if ((++spin & 0x3FFu) == 0) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > kWaitTimeoutNs) __trap();
    }
  }
};
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  SpinGuard g;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    g.tick();
  }
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// one lane of a converged warp (the same lane every time)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- fences
// generic proxy <-> async proxy (TMA, tensor core operand reads), all state spaces
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// global state space only: FENCE.VIEW.ASYNC.G, no MEMBAR (the plain form drags a MEMBAR.ALL.GPU along)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor), 2-D tiles into this CTA
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// the same tile delivered to the same shared-memory offset of every CTA in `mask` of this cluster (one L2 read); each
// destination CTA's own mbarrier (same offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void tma_load_2d_mc(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---------------------------------------------------------------- tcgen05
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1 (sm_100);
// same encoding as gemm_tc.cu::make_desc (validated on B200 by clstm_b200_selftest_gemm)
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFF) >> 4);
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ unsigned make_idesc_f16(int m, int n) {
  return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                        unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same, arriving on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(unsigned bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8 consecutive accumulator columns of this thread's TMEM lane; NO wait (pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld8_nowait(unsigned taddr, unsigned* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int NCOL>
__device__ __forceinline__ void tmem_ld(unsigned taddr, float* v) {   // NCOL multiple of 8, one wait for all loads
  unsigned r[NCOL];
#pragma unroll
  for (int q = 0; q < NCOL / 8; q++) tmem_ld8_nowait(taddr + 8 * q, r + 8 * q);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < NCOL; i++) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- fp16 hi/lo split
// x*scale = hi + lo (+ <= 2^-22 relative): hi = fp16(x*scale), lo = fp16(x*scale - hi).  fp16 has the same 11-bit
// significand as TF32, so hi*hi + hi*lo + lo*hi accumulated in fp32 carries the same ~2^-21 relative error as the 3xTF32
// products of gemm_tc.cu at half the bytes and twice the K per instruction.  `scale` is a power of two that keeps hi well
// inside the fp16 normal range (lo may be subnormal: absolute error <= 2^-25 / scale).  Saturating conversions: a value
// beyond the fp16 range clamps instead of turning into inf.
__device__ __forceinline__ unsigned short f2h_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float h2f(unsigned short h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
__device__ __forceinline__ void split_f16(float xs, unsigned short& hi, unsigned short& lo) {
  hi = f2h_sat(xs);
  lo = f2h_sat(xs - h2f(hi));
}
__device__ __forceinline__ unsigned pack_h2(unsigned short a, unsigned short b) { return (unsigned)a | ((unsigned)b << 16); }

// ---------------------------------------------------------------- cross-CTA step counters in global memory
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Poll with relaxed loads (an acquire load invalidates L1 on every iteration), then ONE acquire fence.  The data the
// counter guards is read through L2 only (TMA, or ld.global.cg), never through a stale L1 line.
template <bool ACQUIRE = true>
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
  SpinGuard g;
  while (ld_relaxed_gpu(p) < target) g.tick();
  if (ACQUIRE) asm volatile("fence.acq_rel.gpu;" ::: "memory");   // (a TMA consumer issues fence.proxy.async instead)
}
// release-increment: every write that happened-before (own writes, and the other threads' writes ordered by the
// preceding CTA barrier) is visible at GPU scope before the counter moves.  Lighter than __threadfence() + atomicAdd,
// which compiles to MEMBAR.SC + an L1 invalidation.
__device__ __forceinline__ void signal_counter(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// ---------------------------------------------------------------- fast gate math (same approximations as lstm.cu)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-kLog2e * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * x)), -1.f); }

}  // namespace tc
}  // namespace cb200
