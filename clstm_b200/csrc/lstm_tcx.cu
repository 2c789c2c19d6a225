// lstm_tcx.cu -- the recurrence of the NPLSTM as a CLUSTER-RESIDENT tensor-core kernel (nhidden 33..480):
// a thread-block cluster owns a group of 16 text lines of one direction for the whole sequence, the recurrent matrix is
// split over the CTAs (32 hidden units = 128 gate rows each; the slice stays in shared memory or -- wide nets -- in TENSOR
// MEMORY for the whole sequence), the per-step products run on tcgen05 (lines on the UMMA N dimension, accumulators in TMEM)
// and h travels between the CTAs through DISTRIBUTED SHARED MEMORY (bulk copies + mbarrier complete_tx) -- no L2 round trip
// inside a step, which is what bounds the lock-step kernels of lstm_tc.cu at ~5 us per step.
//
// Reference semantics (paths relative to /root/reference), identical to lstm.cu / lstm_tc.cu:
//   GenericNPLSTM<SIG,TANH,TANH>::forward   clstm.cc:600-621 (loop body :612-620), forward_lin1 clstm_compute.cc:286
//   GenericNPLSTM::backward                 clstm.cc:622-653 (loop body :629-650), backward_lin1 clstm_compute.cc:296
//
// Geometry.  CS = ceil(nhidden / 32) CTAs per cluster; CTA c owns hidden units 32c .. 32c+31 = gate rows 128c .. 128c+127
// (gate-interleaved, row 4j+g) as ONE UMMA M tile of 128 rows (rows of units >= nhidden are zero).  K index = hidden unit,
// KQ = 32 CS, so the 8 units a warp quadrant of CTA c produces form exactly one 16-byte chunk of the K-major operand row.
//   forward :  pre[128 rows x 16 lines] = R_slice[128 x KQ] (A, resident) * h_{s-1}[16 lines x KQ]^T (B, all-gathered per step)
//   backward:  part[128 outputs x 16 lines] (per 128-output tile) = Rt_slice[KQ x 128 rows] (A, resident) * delta[16 x 128]^T (B, local)
// Operands are fp16 hi/lo pairs (tc_common.cuh::split_f16); B stacks the hi rows (0..15) over the lo rows (16..31), so one
// MMA with N = 32 gives A_hi*[B_hi ; B_lo] and a second one with N = 16 adds A_lo*B_hi: two tcgen05.mma per 16 k.
// A operand: shared memory (SWIZZLE_128B tiles loaded once by TMA) while it fits (KQ <= 256); beyond that the slice lives in
// TMEM (tcgen05.st once, then `tcgen05.mma [d], [a], b-desc`: no shared-memory read of the weights at all) -- forward both
// planes, backward the hi plane (the lo plane stays in shared memory: 512 TMEM columns hold D + one plane of four tiles).
// Threads: 16 epilogue warps (TMEM lane quadrant q = warp & 3 -> 8 units, line quad lg = warp >> 2 -> lines 4lg..4lg+3)
// + 1 loader / MMA warp.  TMEM lane = gate row: the four gates of a unit sit in four adjacent lanes; a 4x4 register transpose
// inside the lane quad (4 shuffles) hands lane g the four gates of line 4lg+g, whose cell state it keeps in a register.
// Exchange: every warp packs its 8 units x 4 lines into 16-byte chunks (8 shuffles) and writes them into a 2 KB staging block
// of its OWN shared memory; the block ([32 rows = 16 lines hi / lo][32 k], SWIZZLE_64B, exactly one K chunk of the B operand)
// then travels with ONE bulk copy per destination CTA (cp.async.bulk shared::cta -> shared::cluster) that completes its bytes
// on a per-source mbarrier of the destination, whose MMA warp starts on a chunk as soon as it has landed.  (The first version
// pushed 16-byte st.async stores: ~3.6 cycles per store and destination, 4900 cycles per step at nhidden 200.)  Two B buffers
// and two staging blocks; safe by data flow (a CTA sends step s only after it received all of step s-1).  The backward
// partial sums travel the same way (2 KB per owner CTA).  Streamed operands (input projection; gates / cell / upstream
// deltas) are prefetched one step ahead into registers.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace cb200 {
namespace {
using namespace tc;

constexpr int kXL = 16;                      // lines per cluster
constexpr int kXEW = 16;                     // epilogue warps
constexpr int kXThreads = 32 * (kXEW + 1);   // + the loader / MMA warp
// Build-time tuning switches (make TCX_FLAGS=n, measured on B200 with tools/tcx_flag_sweep.sh; nhidden 400 x 256 lines, forward / backward ms):
//  0  all-lane barrier waits, the MMA warp issues the bulk copies, MMAs of a step back to back after the last chunk   18.6 / 18.2
//  1  epilogue waits by one lane per warp                                                                            19.9 / 34.3
//  2  MMA-warp waits by one lane                                                                                     22.9 / 32.2
//  4  bulk copies issued by the epilogue warps, one each, behind a named barrier (no longer buildable)               18.0 / 29.9 -> no gain
//  8  MMAs issued chunk by chunk as the chunks arrive (uniform chunk order)                                          22.5 / --
// 16  backward: output tiles innermost (consecutive MMAs to different accumulators)                                   -- / 21.5
// 32  pull form: the CTAs signal "staged", every CTA fetches the peers' blocks with ld.shared::cluster               17.8 / 25.0
// (bits combine).  None of the alternatives beat the plain form: the exchange runs at ~5 B/clk per SM whichever way the bytes
// move (st.async stores, bulk copies, remote loads), so the step is bound by the all-gather itself.
#ifndef CB200_TCX_FLAGS
#define CB200_TCX_FLAGS 0
#endif
constexpr int kXFlags = CB200_TCX_FLAGS;
static_assert((kXFlags & 4) == 0, "variant 4 sent a CTA's own block with a bulk copy (shared::cta -> shared::cluster must target another CTA): kept as a measurement, not buildable");
constexpr int kXDbgStep = 64;                // step whose timeline CTA 0 records when a debug buffer is given
constexpr int kXMaxCS = 15;                  // CTAs per cluster (KQ = 480: D + both A planes fill the 512 TMEM columns)
constexpr float kXScaleH = 16.f, kXScaleR = 16.f, kXScaleD = 256.f;

struct TcxArgs {
  int no, no4, CS, KQ, nks, nkc, nmt;   // hidden units, gate rows, CTAs per cluster, K slots, 16-wide k steps, 64-wide K chunks, 128-output tiles
  int ngroups, d0, ndir, hstride, hoff[2];
  const float* XP[2];
  float* G[2];
  float* C[2];
  float* Hprev[2];
  float* H;
  // backward
  const float* dH;
  float* DG[2];
  // the weight copies (for the TMEM-resident form, which loads them with plain loads)
  const __half *w_hi, *w_lo;
  int nlt;                 // backward, TMEM form: lo-plane tiles that live in tensor memory (the others stay in shared memory)
  long long* dbg;          // optional: timeline of one step of CTA 0 (clock64 stamps)
};

__device__ __forceinline__ unsigned mapa_u32(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// one contiguous block of this CTA's shared memory -> a peer CTA's shared memory; completes `bytes` on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(unsigned rdst, unsigned src, unsigned bytes, unsigned rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst), "r"(src),
               "r"(bytes), "r"(rbar)
               : "memory");
}
// K-major SWIZZLE_64B tile: rows of 64 bytes (32 halfs), 16-byte chunk c of row r at c ^ ((r >> 1) & 3), 8-row groups 512 B apart
__device__ __forceinline__ unsigned long long make_desc64(unsigned saddr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFF) >> 4);
  d |= (unsigned long long)(512 >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;
  return d;
}
__device__ __forceinline__ void sts_v4(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Barriers whose bytes arrive through bulk copies issued by other CTAs are waited on with the plain CTA-scope try_wait of
// tc_common.cuh (as for TMA loads): a cluster-scope acquire compiles to TRYWAIT + CCTL.IVALL, an L1 invalidate per wait that
// cost ~350 cycles per chunk and threw away the prefetched operands.
// One lane polls, the warp follows: a warp-wide try_wait is 32 requests to the barrier unit (measured ~350 cycles per wait of
// the MMA warp, and 512 polling threads per accumulator barrier); __syncwarp orders the other lanes behind the acquire.
__device__ __forceinline__ void mbar_wait_warp(unsigned bar, unsigned parity) {
  if (elect_one()) mbar_wait(bar, parity);
  __syncwarp();
}
// cluster-scope signalling for the pull form of the exchange: a release-arrive on a PEER's barrier, an acquire wait on one's own
__device__ __forceinline__ void mbar_arrive_remote_release(unsigned rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(unsigned bar, unsigned parity) {
  unsigned done = 0;
  SpinGuard g;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    g.tick();
  }
}
__device__ __forceinline__ uint4 ld_cluster_v4(unsigned raddr) {
  uint4 v;
  asm volatile("ld.shared::cluster.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(raddr) : "memory");
  return v;
}
// Pull form: every CTA copies the CS staged 2 KB blocks (block cc from CTA cc's shared memory at `src_off`) into its own buffer
// at dst0 + cc*2048 with 16-byte distributed-shared-memory loads spread over the 512 epilogue threads.
__device__ __forceinline__ void pull_blocks(unsigned dst0, unsigned src_local, int CS, int etid) {
  const int npieces = CS * 128;
  uint4 v[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int idx = etid + 512 * r;
    if (idx < npieces) v[r] = ld_cluster_v4(mapa_u32(src_local + (unsigned)(idx & 127) * 16u, (unsigned)(idx >> 7)));
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int idx = etid + 512 * r;
    if (idx < npieces) sts_v4(dst0 + (unsigned)idx * 16u, v[r].x, v[r].y, v[r].z, v[r].w);
  }
}
__device__ __forceinline__ void mbar_wait_sel(unsigned bar, unsigned parity, bool one_lane) {
  if (one_lane) mbar_wait_warp(bar, parity);
  else mbar_wait(bar, parity);
}
__device__ __forceinline__ void epi_bar_sync() { named_bar_sync(1, 32 * 16); }   // the 16 epilogue warps
// 4x4 transpose inside a quad of lanes: in a[m] = value of THIS lane's gate for line m; out b[k] = gate k of line g
__device__ __forceinline__ void quad_transpose(const float* a, float* b, int g) {
  const bool hi = (g & 2) != 0, lo = (g & 1) != 0;
  const float r0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2);
  const float r1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
  const float own0 = hi ? a[2] : a[0], own1 = hi ? a[3] : a[1];        // my gate, lines 2hi, 2hi+1
  const float ra = __shfl_xor_sync(0xffffffffu, lo ? own0 : own1, 1);   // lane g^1's gate for my line
  const float rb = __shfl_xor_sync(0xffffffffu, lo ? r0 : r1, 1);       // lane g^3's gate for my line
  const float mine = lo ? own1 : own0, par = lo ? r1 : r0;              // gates g and g^2 of line g
  // b[k] = value of gate k: k^g = 0 mine, 1 ra, 2 par, 3 rb  (selects with static k: no local-memory indexing)
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const bool xl = ((k & 1) != 0) != lo, xh = ((k & 2) != 0) != hi;
    b[k] = xh ? (xl ? rb : par) : (xl ? ra : mine);
  }
}
// A operand from tensor memory (K-major, lane = row, 32-bit column = two consecutive k)
__device__ __forceinline__ void mma_f16_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st4(unsigned taddr, uint4 v) {   // 4 consecutive columns of this thread's TMEM lane
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4_nowait(unsigned taddr, unsigned* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_alloc(unsigned dst_smem, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ unsigned pow2_cols(unsigned need) {
  unsigned c = 32;
  while (c < need) c <<= 1;
  return c;
}
__device__ __forceinline__ float ldg_f32(const float* p) { return __ldg(p); }

__device__ __forceinline__ void dbg_stamp(long long* dbg, int slot) {
  if (dbg) dbg[slot] = clock64();
}

// ================================================================================================ forward
// smem (dynamic, 1024-aligned): [A hi: nkc x 16 KB | A lo: nkc x 16 KB] (shared-memory form only; SWIZZLE_128B chunks of 64 k) |
// B buffers 2 x CS x 2 KB (chunk cc of a buffer = the units of CTA cc: [32 rows = 16 lines hi, 16 lines lo][64 B = 32 k],
// SWIZZLE_64B) | staging 2 x 2 KB (this CTA's own chunk of the step, same layout)
// TMEM: D 32 columns | (TMEM form) A hi KQ/2 columns | A lo KQ/2 columns
template <bool A_TMEM>
__global__ void __launch_bounds__(kXThreads, 1)
lstm_tcx_fwd(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, Lines ln, TcxArgs p) {
  extern __shared__ __align__(1024) unsigned char xs[];
  __shared__ __align__(8) unsigned long long bars[2 * kXMaxCS + 6];   // hbar[2][kXMaxCS] (h chunks by source CTA), accbar, abar (weights), stagebar, readybar[2], fullbar
  __shared__ unsigned tmem_base_s;
  __shared__ int lineT[kXL], lineOff[kXL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned CS = cluster_nctarank(), c = cluster_ctarank();
  const int cluster_id = blockIdx.x / (int)CS, nclusters = gridDim.x / (int)CS;
  const int nkc = p.nkc;
  const unsigned smem0 = (smem_u32(xs) + 1023u) & ~1023u;
  const unsigned a_bytes = A_TMEM ? 0u : (unsigned)nkc * 16384u;
  const unsigned a_hi0 = smem0, a_lo0 = smem0 + a_bytes;       // A: chunk kc at +kc*16384 ([128 rows][128 B])
  const unsigned b0 = a_lo0 + a_bytes;                         // B buffer b at b0 + b*bbytes, chunk of CTA cc at +cc*2048
  const unsigned bbytes = CS * 2048u;
  const unsigned stg0 = b0 + 2u * bbytes;                      // staging block s&1 at stg0 + (s&1)*2048
  const unsigned bar0 = smem_u32(&bars[0]);
  const unsigned accbar = bar0 + 8u * (2 * kXMaxCS), abar = accbar + 8u, stagebar = accbar + 16u, readybar0 = accbar + 24u, fullbar = accbar + 40u;
  const bool pull = (kXFlags & 32) != 0;
  const unsigned tcols = pow2_cols(32u + (A_TMEM ? (unsigned)p.KQ : 0u));
  const unsigned acol_hi = 32u, acol_lo = 32u + (unsigned)p.KQ / 2u;
  long long* const dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;

  if (tid == 0) {
    for (int i = 0; i < 2 * kXMaxCS + 2; i++) mbar_init(bar0 + 8u * i, 1);
    mbar_init(stagebar, kXEW);
    mbar_init(readybar0, CS); mbar_init(readybar0 + 8u, CS); mbar_init(fullbar, kXEW);
    mbar_init_fence();
    if (!A_TMEM) { tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); }
  }
  __syncthreads();
  if (warp == kXEW) tmem_alloc(smem_u32(&tmem_base_s), tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;
  cluster_sync_all();                                           // barriers of every CTA initialised before any remote copy
  // this CTA's weight slice of direction d: gate rows [(d*CS + c)*128, +128) of the padded copy (called by all threads)
  unsigned aph = 0;
  auto load_weights = [&](int d) {
    if (!A_TMEM) {
      if (warp == kXEW && elect_one()) {
        mbar_expect_tx(abar, (unsigned)nkc * 2u * 16384u);
        for (int kc = 0; kc < nkc; kc++) {
          tma_load_2d(a_hi0 + kc * 16384, &tmA_hi, kc * 64, (d * (int)CS + (int)c) * 128, abar);
          tma_load_2d(a_lo0 + kc * 16384, &tmA_lo, kc * 64, (d * (int)CS + (int)c) * 128, abar);
        }
      }
      if (warp == kXEW) { mbar_wait(abar, aph); }
      aph ^= 1;
    } else {
      tc_fence_after();
      if (warp < kXEW) {                  // -> tensor memory: lane = gate row, column = k pair; warp lg takes a quarter of K
        const int row = 32 * (warp & 3) + lane, lg = warp >> 2;
        const size_t base = ((size_t)(d * (int)CS + (int)c) * 128 + row) * p.KQ;
        const int kq4 = p.KQ / 4;          // = 8 CS halfs
        const unsigned trow = tmem_d + ((unsigned)(32 * (warp & 3)) << 16);
        for (int k = lg * kq4; k < (lg + 1) * kq4; k += 8) {
          const uint4 vh = *reinterpret_cast<const uint4*>(p.w_hi + base + k);
          const uint4 vl = *reinterpret_cast<const uint4*>(p.w_lo + base + k);
          tmem_st4(trow + acol_hi + (unsigned)k / 2u, vh);
          tmem_st4(trow + acol_lo + (unsigned)k / 2u, vl);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  };

  unsigned hph = 0, accph = 0, stph = 0, fph = 0, rph = 0;      // hph bit b: phase of the chunk barriers of buffer b
  // work items = (line group, direction), longest first (the groups are cut from the length-sorted line order); round r hands
  // item r*nclusters + i to cluster i, odd rounds in reverse (snake), so every cluster gets a similar number of steps
  const int nitems = p.ngroups * p.ndir;
  int cur_d = -1;
  for (int r = 0; r * nclusters < nitems; r++) {
    const int item = r * nclusters + ((r & 1) ? nclusters - 1 - cluster_id : cluster_id);
    if (item >= nitems) continue;
    const int group = item / p.ndir, d = p.d0 + item % p.ndir;
    if (d != cur_d) { load_weights(d); cur_d = d; }
    const int l0 = group * kXL;
    if (tid < kXL) {
      const int li = (l0 + tid < ln.B) ? ln.order[l0 + tid] : -1;
      lineT[tid] = (li >= 0) ? ln.T[li] : 0;
      lineOff[tid] = (li >= 0) ? ln.off[li] : 0;
    }
    __syncthreads();
    const int Tg = lineT[0];

    if (warp == kXEW) {
      // ------------------------------------------------------------------------------------------ copy + MMA issuer
      const unsigned idesc32 = make_idesc_f16(128, 32), idesc16 = make_idesc_f16(128, 16);
      const unsigned long long dbase = make_desc(0);
      auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
      const bool w1 = (kXFlags & 2) != 0, per_chunk = (kXFlags & 8) != 0, epi_copies = (kXFlags & 4) != 0;
      const unsigned cpeer = (unsigned)lane < CS ? (unsigned)lane : 0u;    // lane X < CS sends this CTA's chunk to CTA X
      const unsigned crdst0 = mapa_u32(b0 + c * 2048u, cpeer), crbar0 = mapa_u32(bar0 + 8u * c, cpeer);
      for (int s = 0; s < Tg; s++) {
        const bool rec = dbg && s == kXDbgStep;
        if (s >= 1) {
          const unsigned b = (unsigned)(s - 1) & 1u;            // h_{s-1} sits in buffer (s-1)&1
          const unsigned hb0 = bar0 + 8u * (b * kXMaxCS);
          if (!pull && (unsigned)lane < CS) {                   // (the own chunk is copied locally, see below: a plain arrival keeps its barrier in phase)
            if ((unsigned)lane == c) mbar_arrive(hb0 + 8u * lane);
            else mbar_expect_tx(hb0 + 8u * lane, 2048u);
          }
          __syncwarp();
          const unsigned ph = (hph >> b) & 1u;
          if (pull) {                                           // the epilogue warps have fetched every block of h_{s-1}
            mbar_wait(fullbar, fph);
            fph ^= 1;
            if (rec) dbg_stamp(dbg, 12);
            tc_fence_after();
            if (elect_one()) {
              for (int cc = 0; cc < p.CS; cc++) {
                const unsigned long long bh = make_desc64(b0 + b * bbytes + cc * 2048u);
#pragma unroll
                for (int ks = 0; ks < 2; ks++) {
                  const unsigned kk = 2u * cc + ks;
                  const unsigned acc = (cc > 0 || ks > 0) ? 1u : 0u;
                  if (A_TMEM) {
                    mma_f16_ts(tmem_d, tmem_d + acol_hi + 8u * kk, bh + 2 * ks, idesc32, acc);
                    mma_f16_ts(tmem_d, tmem_d + acol_lo + 8u * kk, bh + 2 * ks, idesc16, 1u);
                  } else {
                    mma_f16(tmem_d, desc_of(a_hi0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc32, acc);
                    mma_f16(tmem_d, desc_of(a_lo0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc16, 1u);
                  }
                }
              }
              mma_commit(accbar);
            }
            __syncwarp();
          } else if (per_chunk) {
            mbar_wait(hb0 + 8u * c, ph);                        // own chunk: every epilogue warp of this CTA has read the accumulator
            tc_fence_after();                                   // of step s-1; then the chunks in UNIFORM order (MMA operands that
            for (int cc = 0; cc < p.CS; cc++) {                 // depend on the CTA rank leave the uniform datapath: ~85 cycles per MMA)
              mbar_wait(hb0 + 8u * cc, ph);
              if (elect_one()) {
                const unsigned long long bh = make_desc64(b0 + b * bbytes + cc * 2048u);
#pragma unroll
                for (int ks = 0; ks < 2; ks++) {
                  const unsigned kk = 2u * cc + ks;
                  const unsigned acc = (cc > 0 || ks > 0) ? 1u : 0u;
                  if (A_TMEM) {
                    mma_f16_ts(tmem_d, tmem_d + acol_hi + 8u * kk, bh + 2 * ks, idesc32, acc);    // R_hi [h_hi ; h_lo]
                    mma_f16_ts(tmem_d, tmem_d + acol_lo + 8u * kk, bh + 2 * ks, idesc16, 1u);     // + R_lo h_hi
                  } else {
                    mma_f16(tmem_d, desc_of(a_hi0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc32, acc);
                    mma_f16(tmem_d, desc_of(a_lo0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc16, 1u);
                  }
                }
                if (cc == p.CS - 1) mma_commit(accbar);
              }
              __syncwarp();
              if (rec && cc == 0) dbg_stamp(dbg, 8);
              if (rec && cc == p.CS - 2) dbg_stamp(dbg, 12);
            }
          } else {                                              // every chunk first, then all MMAs of the step back to back
            for (int i = 0; i < p.CS; i++) {
              mbar_wait_sel(hb0 + 8u * i, ph, w1);
              if (rec && i == 0) dbg_stamp(dbg, 8);
            }
            if (rec) dbg_stamp(dbg, 12);
            tc_fence_after();
            if (elect_one()) {
              for (int cc = 0; cc < p.CS; cc++) {
                const unsigned long long bh = make_desc64(b0 + b * bbytes + cc * 2048u);
#pragma unroll
                for (int ks = 0; ks < 2; ks++) {
                  const unsigned kk = 2u * cc + ks;
                  const unsigned acc = (cc > 0 || ks > 0) ? 1u : 0u;
                  if (A_TMEM) {
                    mma_f16_ts(tmem_d, tmem_d + acol_hi + 8u * kk, bh + 2 * ks, idesc32, acc);
                    mma_f16_ts(tmem_d, tmem_d + acol_lo + 8u * kk, bh + 2 * ks, idesc16, 1u);
                  } else {
                    mma_f16(tmem_d, desc_of(a_hi0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc32, acc);
                    mma_f16(tmem_d, desc_of(a_lo0 + (kk >> 2) * 16384u) + 2 * (kk & 3u), bh + 2 * ks, idesc16, 1u);
                  }
                }
              }
              mma_commit(accbar);
            }
            __syncwarp();
          }
          if (rec) dbg_stamp(dbg, 9);
          hph ^= 1u << b;
        }
        if (!pull && !epi_copies && s + 1 < Tg) {               // h_s is staged: one bulk copy per destination CTA
          mbar_wait_sel(stagebar, stph, w1);
          stph ^= 1;
          const unsigned sb = (unsigned)s & 1u;
          // the own chunk: staging -> own B buffer with plain shared-memory copies (a bulk copy shared::cta -> shared::cluster must
          // target ANOTHER CTA); ordered before this warp's MMAs of the next step by program order + the proxy fence
#pragma unroll
          for (int qq = 0; qq < 4; qq++) {
            const unsigned o = 16u * (unsigned)(lane + 32 * qq);
            unsigned x0, x1, x2, x3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(stg0 + sb * 2048u + o) : "memory");
            sts_v4(b0 + sb * bbytes + c * 2048u + o, x0, x1, x2, x3);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if ((unsigned)lane < CS && (unsigned)lane != c) bulk_copy_to_peer(crdst0 + sb * bbytes, stg0 + sb * 2048u, 2048u, crbar0 + sb * (8u * kXMaxCS));
          __syncwarp();
          if (rec) dbg_stamp(dbg, 11);
        }
      }
    } else {
      // ------------------------------------------------------------------------------------------ epilogue warps
      const int q = warp & 3, lg = warp >> 2;
      const int row = 32 * q + lane;                            // TMEM lane = gate row of this CTA
      const int j = row >> 2, g = row & 3;                      // unit slot, gate
      const int unit = 32 * (int)c + j;
      const bool real = unit < p.no;
      const int grow = 4 * unit + g;
      const int no = p.no, no4 = p.no4;
      const float* __restrict__ XPd = p.XP[d];
      float* __restrict__ Gd = p.G[d];
      float* __restrict__ Cd = p.C[d];
      float* __restrict__ Hpd = p.Hprev[d];
      float* __restrict__ Hd = p.H + p.hoff[d];
      const unsigned taddr = tmem_d + ((unsigned)(32 * q) << 16) + 4u * (unsigned)lg;
      constexpr float inv_scale = 1.0f / (kXScaleH * kXScaleR);
      const float gsc = (g == 3) ? -2.f * kLog2e : -kLog2e;
      // the four lines of this warp: lengths and the element offset of (current column, this gate row)
      int Tl[4];
      long long eo[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        Tl[i] = lineT[4 * lg + i];
        eo[i] = ((long long)lineOff[4 * lg + i] + (d ? Tl[i] - 1 : 0)) * no4 + grow;
      }
      const long long estep = d ? -(long long)no4 : (long long)no4;
      const int myT = lineT[4 * lg + g];                         // the line whose cell state this lane keeps
      long long co = (long long)lineOff[4 * lg + g] + (d ? myT - 1 : 0);
      const int cstep = d ? -1 : 1;
      float cst = 0.f;
      // staging: lanes 0..7 hold the warp's eight 16-byte chunks: (line ll = lane & 3, plane = lane >> 2), 8 units of quadrant q
      const int ll = lane & 3, plane = (lane >> 2) & 1;
      const unsigned xrow = (unsigned)(4 * lg + ll + 16 * plane);
      const unsigned xoff = xrow * 64u + ((((unsigned)q) ^ ((xrow >> 1) & 3u)) << 4);
      const bool rec0 = dbg && warp == 0 && lane == 0;
      const unsigned peer = (unsigned)warp < CS ? (unsigned)warp : 0u;
      const unsigned rdst0 = mapa_u32(b0 + c * 2048u, peer), rbar0 = mapa_u32(bar0 + 8u * c, peer);
      float xp[4];
#pragma unroll
      for (int i = 0; i < 4; i++) xp[i] = (real && 0 < Tl[i]) ? ldg_f32(XPd + eo[i]) : 0.f;
      for (int s = 0; s < Tg; s++) {
        const bool rec = rec0 && s == kXDbgStep;
        float xpn[4];                                           // next step's input projection: in flight during this step
#pragma unroll
        for (int i = 0; i < 4; i++) xpn[i] = (real && s + 1 < Tl[i]) ? ldg_f32(XPd + eo[i] + estep) : 0.f;
        float act[4];
        if (s > 0) {
          mbar_wait_sel(accbar, accph, (kXFlags & 1) != 0);
          accph ^= 1;
          if (rec) dbg_stamp(dbg, 0);
          tc_fence_after();
          unsigned ra[4], rb[4];
          tmem_ld4_nowait(taddr, ra);
          tmem_ld4_nowait(taddr + 16u, rb);
          tmem_wait_ld();
          if (rec) dbg_stamp(dbg, 1);
#pragma unroll
          for (int i = 0; i < 4; i++) act[i] = fmaf(__uint_as_float(ra[i]) + __uint_as_float(rb[i]), inv_scale, xp[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) act[i] = xp[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {                            // forward_full1 clstm.cc:614-617: sigmoid for gi, gf, go; tanh(x) = 2 sigmoid(2x) - 1 for ci
          const float y = rcp_approx(1.0f + ex2_approx(gsc * act[i]));
          act[i] = (g == 3) ? fmaf(2.f, y, -1.f) : y;
        }
        float gt[4];
        quad_transpose(act, gt, g);                             // gi, gf, go, ci of line 4lg+g
        const bool on = real && s < myT;
        float hh = 0.f;
        if (on) {
          cst = fmaf(gt[1], cst, gt[3] * gt[0]);                // forward_statemem clstm_compute.cc:504-508
          hh = tanh_fast(cst) * gt[2];                          // forward_nonlingate :530-537
        }
        if (s + 1 < Tg) {
          unsigned short h16, l16;
          split_f16(hh * kXScaleH, h16, l16);
          const unsigned hp = pack_h2(h16, l16);
          // this warp's 8 units x 4 lines as 16-byte chunks: lane (line ll, plane) collects unit k's value from lane 4k + ll
          unsigned v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = __shfl_sync(0xffffffffu, hp, 4 * k + ll);
          unsigned w4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const unsigned x0 = plane ? (v[2 * e] >> 16) : (v[2 * e] & 0xffffu);
            const unsigned x1 = plane ? (v[2 * e + 1] >> 16) : (v[2 * e + 1] & 0xffffu);
            w4[e] = x0 | (x1 << 16);
          }
          if (rec) dbg_stamp(dbg, 2);
          tc_fence_before();                                    // (the accumulator has been read: tcgen05.wait::ld above)
          const unsigned sb = (unsigned)s & 1u;
          if (lane < 8) sts_v4(stg0 + sb * 2048u + xoff, w4[0], w4[1], w4[2], w4[3]);
          if (pull) {
            epi_bar_sync();                                     // h_s is staged here: tell every CTA of the cluster
            if (warp == 0 && (unsigned)lane < CS) mbar_arrive_remote_release(mapa_u32(readybar0 + 8u * sb, (unsigned)lane));
            if (rec) dbg_stamp(dbg, 3);
          } else
          fence_proxy_async_smem();                             // generic-proxy stores -> the bulk copies' reads
          if (pull) {
          } else if (kXFlags & 4) {
            epi_bar_sync();                                     // h_s is staged: warp X < CS sends the block to CTA X (one copy each)
            if (rec) dbg_stamp(dbg, 3);
            if ((unsigned)warp < CS && elect_one())
              bulk_copy_to_peer(rdst0 + sb * bbytes, stg0 + sb * 2048u, 2048u, rbar0 + sb * (8u * kXMaxCS));
            if (rec) dbg_stamp(dbg, 10);
          } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(stagebar);               // the MMA warp sends the block once all 16 warps have staged
            if (rec) dbg_stamp(dbg, 3);
          }
        }
        // ---- stash for the backward pass and the dense products
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (real && s < Tl[i]) Gd[eo[i]] = act[i];
        if (on) {
          Cd[co * no + unit] = cst;
          Hd[co * p.hstride + unit] = hh;
          if (s + 1 < myT) Hpd[(co + cstep) * no + unit] = hh;
          if (s == 0) Hpd[co * no + unit] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { eo[i] += estep; xp[i] = xpn[i]; }
        co += cstep;
        if (rec) dbg_stamp(dbg, 4);
        if (pull && s + 1 < Tg) {                               // every CTA has staged h_s: fetch the CS blocks into B buffer s&1
          const unsigned sb = (unsigned)s & 1u;
          if (warp == 0) mbar_wait_acq_cluster(readybar0 + 8u * sb, (rph >> sb) & 1u);
          rph ^= 1u << sb;
          epi_bar_sync();
          if (rec) dbg_stamp(dbg, 10);
          pull_blocks(b0 + sb * bbytes, stg0 + sb * 2048u, p.CS, tid);
          fence_proxy_async_smem();                             // generic-proxy stores -> tensor-core reads
          __syncwarp();
          if (lane == 0) mbar_arrive(fullbar);
          if (rec) dbg_stamp(dbg, 11);
        }
      }
    }
    __syncthreads();
    cluster_sync_all();        // nobody may send the next group's h_0 into a buffer a slower CTA is still multiplying
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kXEW) {
    __syncwarp();
    tmem_dealloc(tmem_d, tcols);
  }
}


// ================================================================================================ backward
// smem: A (Rt slice) chunks (mt, kc) of [128 output rows][64 gate rows], SWIZZLE_128B: hi plane then lo plane (shared-memory
// form), or only the lo tiles nlt .. nmt-1 (TMEM form) | B = delta tile [2 kc][32 rows][128 B] | reduce buffers
// [2][CS src][32 slots][16 lines] fp32 | staging [2][CS owners][32 slots][16 lines] fp32 (16-byte chunks XOR-swizzled by slot)
// TMEM: D nmt x 32 columns | (TMEM form) A hi nmt x 64 columns | A lo nlt x 64 columns
template <bool A_TMEM>
__global__ void __launch_bounds__(kXThreads, 1)
lstm_tcx_bwd(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, Lines ln, TcxArgs p) {
  extern __shared__ __align__(1024) unsigned char xs[];
  __shared__ __align__(8) unsigned long long bars[8];          // pbar[2] (partial sums), accbar, abar (weights), bbar (delta tile), stagebar, readybar[2]
  __shared__ unsigned tmem_base_s;
  __shared__ int lineT[kXL], lineOff[kXL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned CS = cluster_nctarank(), c = cluster_ctarank();
  const int cluster_id = blockIdx.x / (int)CS, nclusters = gridDim.x / (int)CS;
  const int nmt = p.nmt;                                       // 128-row tiles of output slots (the last one may reach beyond KQ)
  const int nlt = A_TMEM ? p.nlt : 0;                          // lo tiles in tensor memory
  const unsigned smem0 = (smem_u32(xs) + 1023u) & ~1023u;
  const unsigned a_hi0 = smem0;                                // chunk (mt, kc) at +(mt*2+kc)*16384 (shared-memory form)
  const unsigned a_lo0 = smem0 + (A_TMEM ? 0u : (unsigned)nmt * 32768u);   // lo tile mt at +((mt - nlt)*2 + kc)*16384
  const unsigned b0 = a_lo0 + (unsigned)(nmt - nlt) * 32768u;  // delta tile: chunk kc at +kc*4096
  const unsigned r0 = b0 + 8192u;                              // reduce buffers
  const unsigned rbytes = CS * 2048u;
  const unsigned stg0 = r0 + 2u * rbytes;                      // staging
  const unsigned bar0 = smem_u32(&bars[0]);
  const unsigned pbar0 = bar0, pbar1 = bar0 + 8, accbar = bar0 + 16, abar = bar0 + 24, bbar = bar0 + 32, stagebar = bar0 + 40, readybar0 = bar0 + 48;
  const bool pull = (kXFlags & 32) != 0;
  const unsigned tcols = pow2_cols((unsigned)nmt * (A_TMEM ? 96u : 32u) + (unsigned)nlt * 64u);
  const unsigned acol_hi = 32u * (unsigned)nmt, acol_lo = 96u * (unsigned)nmt;
  long long* const dbg = (p.dbg && blockIdx.x == 0) ? p.dbg + 32 : nullptr;

  if (tid == 0) {
    mbar_init(pbar0, 2); mbar_init(pbar1, 2); mbar_init(accbar, 1); mbar_init(abar, 1); mbar_init(bbar, (kXFlags & 1) ? kXEW : 32 * kXEW); mbar_init(stagebar, kXEW); mbar_init(readybar0, CS); mbar_init(readybar0 + 8u, CS);
    mbar_init_fence();
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
  }
  for (unsigned i = tid; i < 8192 / 16; i += blockDim.x) sts_v4(b0 + 16 * i, 0u, 0u, 0u, 0u);
  __syncthreads();
  const unsigned a_loads = (unsigned)(A_TMEM ? (nmt - nlt) : 2 * nmt) * 2u;   // 16 KB TMA boxes
  if (warp == kXEW) tmem_alloc(smem_u32(&tmem_base_s), tcols);
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;
  cluster_sync_all();
  // Rt slice of this CTA for direction d: rows = output slots k', columns = this CTA's 128 gate rows (called by all threads)
  unsigned aph = 0;
  auto load_weights = [&](int d) {
    if (a_loads) {
      if (warp == kXEW && elect_one()) {
        mbar_expect_tx(abar, a_loads * 16384u);
        for (int mt = 0; mt < nmt; mt++)
          for (int kc = 0; kc < 2; kc++) {
            const int rw = (d * (int)CS + (int)c) * p.KQ + mt * 128;
            if (!A_TMEM) tma_load_2d(a_hi0 + (mt * 2 + kc) * 16384, &tmA_hi, kc * 64, rw, abar);
            if (mt >= nlt) tma_load_2d(a_lo0 + ((mt - nlt) * 2 + kc) * 16384, &tmA_lo, kc * 64, rw, abar);
          }
      }
      if (warp == kXEW) mbar_wait(abar, aph);
      aph ^= 1;
    }
    if (A_TMEM) {
      tc_fence_after();
      if (warp < kXEW) {                  // -> tensor memory: lane = output slot of the tile, column = gate-row pair
        const int rowl = 32 * (warp & 3) + lane, lg = warp >> 2;
        const unsigned trow = tmem_d + ((unsigned)(32 * (warp & 3)) << 16);
        for (int mt = 0; mt < nmt; mt++) {
          const int slot = mt * 128 + rowl;
          const size_t base = ((size_t)(d * (int)CS + (int)c) * p.KQ + slot) * 128;
#pragma unroll
          for (int e = 0; e < 4; e++) {   // this warp's quarter of the 128 gate rows: 32 halfs = 4 x 16 bytes
            const int k = 32 * lg + 8 * e;
            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = make_uint4(0u, 0u, 0u, 0u);
            if (slot < p.KQ) {
              vh = *reinterpret_cast<const uint4*>(p.w_hi + base + k);
              if (mt < nlt) vl = *reinterpret_cast<const uint4*>(p.w_lo + base + k);
            }
            tmem_st4(trow + acol_hi + 64u * (unsigned)mt + (unsigned)k / 2u, vh);
            if (mt < nlt) tmem_st4(trow + acol_lo + 64u * (unsigned)mt + (unsigned)k / 2u, vl);
          }
        }
        tmem_wait_st();
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  };
  unsigned pph0 = 0, pph1 = 0, accph = 0, bph = 0, stph = 0, rph = 0;
  const unsigned pbytes = CS * 2048u;                           // bytes a CTA receives per step (= one reduce buffer)
  const int nitems = p.ngroups * p.ndir;                        // work items as in the forward kernel
  int cur_d = -1;
  for (int r = 0; r * nclusters < nitems; r++) {
    const int item = r * nclusters + ((r & 1) ? nclusters - 1 - cluster_id : cluster_id);
    if (item >= nitems) continue;
    const int group = item / p.ndir, d = p.d0 + item % p.ndir;
    if (d != cur_d) { load_weights(d); cur_d = d; }
    const int l0 = group * kXL;
    if (tid < kXL) {
      const int li = (l0 + tid < ln.B) ? ln.order[l0 + tid] : -1;
      lineT[tid] = (li >= 0) ? ln.T[li] : 0;
      lineOff[tid] = (li >= 0) ? ln.off[li] : 0;
    }
    __syncthreads();
    const int Tg = lineT[0];

    if (warp == kXEW) {
      // ------------------------------------------------------------------------------------------ MMA + copy issuer
      const unsigned idesc32 = make_idesc_f16(128, 32), idesc16 = make_idesc_f16(128, 16);
      const unsigned long long dbase = make_desc(0);
      auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
      const bool w1 = (kXFlags & 2) != 0, epi_copies = (kXFlags & 4) != 0, tiles_inner = (kXFlags & 16) != 0;
      const unsigned cpeer = (unsigned)lane < CS ? (unsigned)lane : 0u;    // lane X < CS sends owner X's partial sums to CTA X
      const unsigned crdst0 = mapa_u32(r0 + c * 2048u, cpeer), crbar0 = mapa_u32(pbar0, cpeer);
      for (int it = 0; it + 1 < Tg; it++) {
        const bool rec = dbg && it == kXDbgStep;
        mbar_wait_sel(bbar, bph, w1);                           // the deltas of this step are in shared memory
        bph ^= 1;
        tc_fence_after();
        if (elect_one()) {
          if (tiles_inner) {                                    // consecutive MMAs go to different accumulators
            for (int kc = 0; kc < 2; kc++) {
              const unsigned long long bh = desc_of(b0 + kc * 4096);
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                const unsigned acc = (kc > 0 || ks > 0) ? 1u : 0u;
                for (int mt = 0; mt < nmt; mt++) {
                  const unsigned kcol = 64u * mt + 8u * (4 * kc + ks);
                  if (A_TMEM) mma_f16_ts(tmem_d + 32 * mt, tmem_d + acol_hi + kcol, bh + 2 * ks, idesc32, acc);
                  else mma_f16(tmem_d + 32 * mt, desc_of(a_hi0 + (mt * 2 + kc) * 16384) + 2 * ks, bh + 2 * ks, idesc32, acc);
                  if (mt < nlt) mma_f16_ts(tmem_d + 32 * mt, tmem_d + acol_lo + kcol, bh + 2 * ks, idesc16, 1u);
                  else mma_f16(tmem_d + 32 * mt, desc_of(a_lo0 + ((mt - nlt) * 2 + kc) * 16384) + 2 * ks, bh + 2 * ks, idesc16, 1u);
                }
              }
            }
          } else {
            for (int mt = 0; mt < nmt; mt++)
              for (int kc = 0; kc < 2; kc++) {
                const unsigned long long bh = desc_of(b0 + kc * 4096);
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                  const unsigned acc = (kc > 0 || ks > 0) ? 1u : 0u;
                  const unsigned kcol = 64u * mt + 8u * (4 * kc + ks);
                  if (A_TMEM) mma_f16_ts(tmem_d + 32 * mt, tmem_d + acol_hi + kcol, bh + 2 * ks, idesc32, acc);               // Rt_hi [d_hi ; d_lo]
                  else mma_f16(tmem_d + 32 * mt, desc_of(a_hi0 + (mt * 2 + kc) * 16384) + 2 * ks, bh + 2 * ks, idesc32, acc);
                  if (mt < nlt) mma_f16_ts(tmem_d + 32 * mt, tmem_d + acol_lo + kcol, bh + 2 * ks, idesc16, 1u);               // + Rt_lo d_hi
                  else mma_f16(tmem_d + 32 * mt, desc_of(a_lo0 + ((mt - nlt) * 2 + kc) * 16384) + 2 * ks, bh + 2 * ks, idesc16, 1u);
                }
              }
          }
          mma_commit(accbar);
        }
        __syncwarp();
        if (rec) dbg_stamp(dbg, 9);
        if (!pull && !epi_copies) {
          mbar_wait_sel(stagebar, stph, w1);                    // the partial sums are staged: one bulk copy per owner CTA
          stph ^= 1;
          const unsigned sb = (unsigned)it & 1u;
          // the block of my own slots stays in the staging buffer (the reduction reads it there: a bulk copy must target ANOTHER CTA);
          // this arrival publishes it to the epilogue warps together with the remote blocks
          if (lane == 0) mbar_arrive(sb ? pbar1 : pbar0);
          if ((unsigned)lane < CS && (unsigned)lane != c)
            bulk_copy_to_peer(crdst0 + sb * rbytes, stg0 + sb * rbytes + (unsigned)lane * 2048u, 2048u, crbar0 + sb * 8u);
          __syncwarp();
          if (rec) dbg_stamp(dbg, 11);
        }
      }
    } else {
      // ------------------------------------------------------------------------------------------ epilogue warps
      const int q = warp & 3, lg = warp >> 2;
      const int row = 32 * q + lane;
      const int j = row >> 2, g = row & 3;
      const int unit = 32 * (int)c + j;
      const bool real = unit < p.no;
      const int grow = 4 * unit + g;
      const int no = p.no, no4 = p.no4;
      const float* __restrict__ Gd = p.G[d];
      const float* __restrict__ Cd = p.C[d];
      const float* __restrict__ dHd = p.dH + p.hoff[d];
      float* __restrict__ DGd = p.DG[d];
      const unsigned taddr = tmem_d + ((unsigned)(32 * q) << 16) + 4u * (unsigned)lg;
      constexpr float inv_scale = 1.0f / (kXScaleD * kXScaleR);
      // the four lines of this warp; forward step fs of line l sits in column off + (d ? T-1-fs : fs)
      const int fs0 = Tg - 1;
      int Tl[4];
      long long eo[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        Tl[i] = lineT[4 * lg + i];
        eo[i] = ((long long)lineOff[4 * lg + i] + (d ? Tl[i] - 1 - fs0 : fs0)) * no4 + grow;
      }
      const long long estep = d ? (long long)no4 : -(long long)no4;   // fs -> fs - 1
      const int myT = lineT[4 * lg + g];
      long long co = (long long)lineOff[4 * lg + g] + (d ? myT - 1 - fs0 : fs0);
      const int cstep = d ? 1 : -1;
      float dcc = 0.f;                                          // carried cell derivative of (unit j, line 4lg+g)
      // where this thread's deltas (gate row r' = row) go in the delta tile
      const unsigned boff = (unsigned)(row >> 6) * 4096u + 2u * (unsigned)(row & 7);
      const unsigned bchunk = (unsigned)((row & 63) >> 3);
      // reduce buffer: element (source sc, slot j, line 4lg+g); staging: (owner, slot lane, lines 4lg..4lg+3); chunks swizzled by slot
      const unsigned red_off = (unsigned)j * 64u + ((((unsigned)lg) ^ (((unsigned)j >> 1) & 3u)) << 4) + (unsigned)g * 4u;
      const unsigned stg_off = (unsigned)lane * 64u + ((((unsigned)lg) ^ (((unsigned)lane >> 1) & 3u)) << 4);
      const bool rec0 = dbg && warp == 0 && lane == 0;
      const unsigned peer = (unsigned)warp < CS ? (unsigned)warp : 0u;
      const unsigned rdst0 = mapa_u32(r0 + c * 2048u, peer), rbar0 = mapa_u32(pbar0, peer);
      // operands of the first step
      float gact[4], cc = 0.f, cp = 0.f, dh = 0.f;
#pragma unroll
      for (int i = 0; i < 4; i++) gact[i] = (real && fs0 < Tl[i]) ? ldg_f32(Gd + eo[i]) : 0.f;
      if (real && fs0 < myT) {
        cc = ldg_f32(Cd + co * no + unit);
        if (fs0 > 0) cp = ldg_f32(Cd + (co + cstep) * no + unit);
        dh = ldg_f32(dHd + co * p.hstride + unit);
      }
      for (int it = 0; it < Tg; it++) {
        const int fs = Tg - 1 - it;
        const bool rec = rec0 && it == kXDbgStep;
        // ---- next step's operands: in flight during this step (they do not depend on the exchange)
        float gactn[4], ccn = 0.f, cpn = 0.f, dhn = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) gactn[i] = (real && fs >= 1 && fs - 1 < Tl[i]) ? ldg_f32(Gd + eo[i] + estep) : 0.f;
        if (real && fs >= 1 && fs - 1 < myT) {
          const long long cn = co + cstep;
          ccn = ldg_f32(Cd + cn * no + unit);
          if (fs - 1 > 0) cpn = ldg_f32(Cd + (cn + cstep) * no + unit);
          dhn = ldg_f32(dHd + cn * p.hstride + unit);
        }
        // ---- recurrent part of dh: the CS partial products of the previous step, fixed order
        if (it > 0) {
          const unsigned b = (unsigned)(it - 1) & 1u;
          const unsigned pb = b ? pbar1 : pbar0;
          if (pull) {                                           // every CTA has staged its partial sums: fetch the blocks of MY slots
            if (warp == 0) mbar_wait_acq_cluster(readybar0 + 8u * b, (rph >> b) & 1u);
            rph ^= 1u << b;
            epi_bar_sync();
            pull_blocks(r0 + b * rbytes, stg0 + b * rbytes + c * 2048u, p.CS, tid);
            epi_bar_sync();
          } else {
            if (tid == 0) mbar_expect_tx(pb, pbytes - 2048u);   // (CS - 1 remote blocks; the own one is read from the staging buffer)
            mbar_wait_sel(pb, b ? pph1 : pph0, (kXFlags & 1) != 0);
            if (b) pph1 ^= 1; else pph0 ^= 1;
          }
          if (rec) dbg_stamp(dbg, 0);
          if (real && fs < myT - 1) {
            float r = 0.f;
            const unsigned ra = r0 + b * rbytes + red_off;
            const unsigned own = (kXFlags & 32) ? ra + c * 2048u : stg0 + b * rbytes + c * 2048u + red_off;
            for (unsigned sc = 0; sc < CS; sc++) {
              float x;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(sc == c ? own : ra + sc * 2048u) : "memory");
              r += x;
            }
            dh += r;
          }
        }
        // ---- pointwise (backward_nonlingate / statemem / nonlin0, clstm_compute.cc:539-547, 509-515, 231-267)
        float gt[4], dq4[4], dl[4];
        quad_transpose(gact, gt, g);                            // gi, gf, go, ci of line 4lg+g
        const bool on = real && fs < myT;
        if (on) {
          const float th = tanh_fast(cc);
          const float dgo = th * dh;
          const float dc = fmaf(1.f - th * th, gt[2] * dh, dcc);
          float dgf = 0.f, carry = 0.f;
          if (fs > 0) { dgf = dc * cp; carry = dc * gt[1]; }
          dcc = carry;
          dq4[0] = gt[0] * (1.f - gt[0]) * (dc * gt[3]);
          dq4[1] = gt[1] * (1.f - gt[1]) * dgf;
          dq4[2] = gt[2] * (1.f - gt[2]) * dgo;
          dq4[3] = (1.f - gt[3] * gt[3]) * (dc * gt[0]);
        } else { dq4[0] = dq4[1] = dq4[2] = dq4[3] = 0.f; }
        quad_transpose(dq4, dl, g);                             // back: this lane's gate for lines 4lg..4lg+3
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (real && fs < Tl[i]) DGd[eo[i]] = dl[i];
        if (fs == 0) break;
        // ---- delta tile (B operand): row = line (+16 for the lo plane), k = this thread's gate row
#pragma unroll
        for (int i = 0; i < 4; i++) {
          unsigned short h16, l16;
          split_f16(dl[i] * kXScaleD, h16, l16);
          const unsigned rh = (unsigned)(4 * lg + i), rl = rh + 16u;
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(b0 + boff + rh * 128u + ((bchunk ^ (rh & 7u)) << 4)), "h"(h16) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(b0 + boff + rl * 128u + ((bchunk ^ (rl & 7u)) << 4)), "h"(l16) : "memory");
        }
        fence_proxy_async_smem();
        if (kXFlags & 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bbar);
        } else mbar_arrive(bbar);
        if (rec) dbg_stamp(dbg, 1);
        // ---- partial products of this CTA's gate rows for ALL output slots: stage each owner's slots x lines block
        mbar_wait_sel(accbar, accph, (kXFlags & 1) != 0);
        accph ^= 1;
        if (rec) dbg_stamp(dbg, 2);
        tc_fence_after();
        for (int mt = 0; mt < nmt; mt++) {
          unsigned ra[4], rb[4];
          tmem_ld4_nowait(taddr + 32u * mt, ra);
          tmem_ld4_nowait(taddr + 32u * mt + 16u, rb);
          tmem_wait_ld();
          const unsigned owner = (unsigned)(4 * mt + q);        // slots 128 mt + 32 q .. +31 belong to this CTA of the cluster
          if (owner < CS)
            sts_v4(stg0 + ((unsigned)it & 1u) * rbytes + owner * 2048u + stg_off,
                   __float_as_uint((__uint_as_float(ra[0]) + __uint_as_float(rb[0])) * inv_scale),
                   __float_as_uint((__uint_as_float(ra[1]) + __uint_as_float(rb[1])) * inv_scale),
                   __float_as_uint((__uint_as_float(ra[2]) + __uint_as_float(rb[2])) * inv_scale),
                   __float_as_uint((__uint_as_float(ra[3]) + __uint_as_float(rb[3])) * inv_scale));
        }
        tc_fence_before();
        if (pull) {
          epi_bar_sync();                                       // staged: tell every CTA of the cluster
          if (warp == 0 && (unsigned)lane < CS) mbar_arrive_remote_release(mapa_u32(readybar0 + 8u * ((unsigned)it & 1u), (unsigned)lane));
          if (rec) dbg_stamp(dbg, 3);
        } else
        fence_proxy_async_smem();
        if (pull) {
        } else if (kXFlags & 4) {
          epi_bar_sync();                                       // staged: warp X < CS sends the block of owner X (slot c of its reduce buffer)
          if (rec) dbg_stamp(dbg, 3);
          if ((unsigned)warp < CS && elect_one()) {
            const unsigned sb = (unsigned)it & 1u;
            bulk_copy_to_peer(rdst0 + sb * rbytes, stg0 + sb * rbytes + (unsigned)warp * 2048u, 2048u, rbar0 + sb * 8u);
          }
          if (rec) dbg_stamp(dbg, 10);
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(stagebar);
          if (rec) dbg_stamp(dbg, 3);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { eo[i] += estep; gact[i] = gactn[i]; }
        co += cstep;
        cc = ccn; cp = cpn; dh = dhn;
      }
    }
    __syncthreads();
    cluster_sync_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kXEW) {
    __syncwarp();
    tmem_dealloc(tmem_d, tcols);
  }
}

// ================================================================================================ weight layouts
// R [4no x no] fp32 -> fp16 hi/lo copies.  Forward: the matrix padded to [CS*128 rows][KQ] (row = gate row, column = unit).
// Backward: for CTA c a [KQ output slots][128 gate rows of c] matrix, element (k', r') = R[128c + r'][k'].  Padding stays
// zero from the allocation.
__global__ void lstm_tcx_split_kernel(const float* __restrict__ R, int no, int KQ, __half* __restrict__ a_hi,
                                      __half* __restrict__ a_lo, __half* __restrict__ t_hi, __half* __restrict__ t_lo) {
  const size_t total = (size_t)4 * no * no;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / no), k = (int)(i % no);
    unsigned short hi, lo;
    split_f16(R[i] * kXScaleR, hi, lo);
    const size_t o = (size_t)r * KQ + k;
    reinterpret_cast<unsigned short*>(a_hi)[o] = hi;
    reinterpret_cast<unsigned short*>(a_lo)[o] = lo;
    const size_t ot = ((size_t)(r >> 7) * KQ + k) * 128 + (r & 127);
    reinterpret_cast<unsigned short*>(t_hi)[ot] = hi;
    reinterpret_cast<unsigned short*>(t_lo)[ot] = lo;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_x = nullptr;
int load_encode_x() {
  if (g_encode_x) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
  g_encode_x = (EncodeTiledFn)fn;
  return 0;
}
int make_map_x(CUtensorMap* m, void* base, size_t rows, size_t cols, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return g_encode_x(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : 1;
}
}  // namespace

struct LstmTcxPlan {
  int no = 0, num_sms = 148, CS = 0, KQ = 0, nks = 0, nkc = 0, nmt = 0;
  bool tmem = false;                             // weight slice in tensor memory (forward: both planes, backward: the hi plane)
  __half *a_hi = nullptr, *a_lo = nullptr;       // forward weight slices [2 dirs][CS*128][KQ]
  __half *t_hi = nullptr, *t_lo = nullptr;       // backward weight slices [2 dirs][CS][KQ][128]
  bool stale[2] = {true, true};
  CUtensorMap tmA_hi, tmA_lo, tmT_hi, tmT_lo;
  int max_clusters = 0;
  int nlt = 0;                                   // backward, TMEM form: lo tiles that fit into tensor memory next to D and the hi plane
  long long* dbg = nullptr;                      // 64 clock stamps (forward 0..31, backward 32..63) when CLSTM_B200_TC_DBG is set
  long long dbg_host[64] = {0};
  char err[256] = {0};
};

bool lstm_tcx_supported(int no) {
  const int cs = (no + 31) / 32;
  return cs >= 2 && cs <= kXMaxCS;
}
void lstm_tcx_destroy(LstmTcxPlan* p) {
  if (!p) return;
  cudaFree(p->a_hi); cudaFree(p->a_lo); cudaFree(p->t_hi); cudaFree(p->t_lo); cudaFree(p->dbg);
  delete p;
}
const char* lstm_tcx_error(const LstmTcxPlan* p) { return p ? p->err : "no plan"; }
void lstm_tcx_mark_stale(LstmTcxPlan* p) { if (p) p->stale[0] = p->stale[1] = true; }
// timeline of step kXDbgStep of CTA 0 (after a synchronize): [0..15] forward, [32..47] backward; nullptr without CLSTM_B200_TC_DBG
const long long* lstm_tcx_debug(LstmTcxPlan* p) {
  if (!p || !p->dbg) return nullptr;
  cudaMemcpy(p->dbg_host, p->dbg, sizeof p->dbg_host, cudaMemcpyDeviceToHost);
  return p->dbg_host;
}

namespace {
size_t tcx_fwd_smem(const LstmTcxPlan* p) { return (p->tmem ? 0 : (size_t)p->nkc * 2 * 16384) + (size_t)2 * p->CS * 2048 + 2 * 2048 + 1024; }
size_t tcx_bwd_smem(const LstmTcxPlan* p) {
  const size_t a = p->tmem ? (size_t)(p->nmt - p->nlt) * 32768 : (size_t)p->nmt * 65536;
  return a + 8192 + (size_t)4 * p->CS * 2048 + 1024;
}
int ensure_split_x(LstmTcxPlan* p, cudaStream_t st, const float* const R[2], int d0, int ndir) {
  for (int d = d0; d < d0 + ndir; d++) {
    if (!p->stale[d]) continue;
    const size_t off = (size_t)d * p->CS * 128 * p->KQ;
    const size_t total = (size_t)4 * p->no * p->no;
    const int nb = (int)std::min<size_t>((total + 255) / 256, (size_t)p->num_sms * 8);
    lstm_tcx_split_kernel<<<nb, 256, 0, st>>>(R[d], p->no, p->KQ, p->a_hi + off, p->a_lo + off, p->t_hi + off, p->t_lo + off);
    p->stale[d] = false;
  }
  return (int)cudaGetLastError();
}
template <class K>
cudaError_t tcx_attrs(K kern, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}
void tcx_launch_cfg(const LstmTcxPlan* p, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int nclusters, size_t smem, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(nclusters * p->CS); cfg.blockDim = dim3(kXThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
}
void tcx_fill(const LstmTcxPlan* p, TcxArgs& x, int B, int d0, int ndir, int hstride, const int* hoff) {
  x.no = p->no; x.no4 = 4 * p->no; x.CS = p->CS; x.KQ = p->KQ; x.nks = p->nks; x.nkc = p->nkc; x.nmt = p->nmt;
  x.ngroups = (B + kXL - 1) / kXL; x.d0 = d0; x.ndir = ndir; x.hstride = hstride; x.hoff[0] = hoff[0]; x.hoff[1] = hoff[1];
  x.nlt = p->nlt; x.dbg = p->dbg;
}
}

LstmTcxPlan* lstm_tcx_create(int no, int num_sms) {
  if (!lstm_tcx_supported(no) || load_encode_x() != 0) return nullptr;
  auto* p = new LstmTcxPlan;
  p->no = no; p->num_sms = num_sms;
  p->CS = (no + 31) / 32; p->KQ = 32 * p->CS; p->nks = p->KQ / 16; p->nkc = (p->KQ + 63) / 64; p->nmt = (p->KQ + 127) / 128;
  // weights in tensor memory: measured faster at every width (nhidden 200, 128 lines: 4.9 + 4.2 ms against 5.7 + 5.5 ms with
  // the slice in shared memory, whose 128 x 16 A tile costs 32 cycles of shared-memory bandwidth per MMA); the shared-memory
  // form stays selectable for A/B runs (CLSTM_B200_TCX_TMEM=0) where it fits
  p->tmem = true;
  if (const char* e = getenv("CLSTM_B200_TCX_TMEM")) p->tmem = (atoi(e) != 0) || p->KQ > 256;
  p->nlt = p->tmem ? std::min(p->nmt, (512 - 96 * p->nmt) / 64) : 0;
  if (getenv("CLSTM_B200_TC_DBG") && cudaMalloc((void**)&p->dbg, sizeof p->dbg_host) == cudaSuccess) cudaMemset(p->dbg, 0, sizeof p->dbg_host);
  const size_t elems = (size_t)2 * p->CS * 128 * p->KQ;
  bool ok = cudaMalloc((void**)&p->a_hi, elems * 2) == cudaSuccess && cudaMalloc((void**)&p->a_lo, elems * 2) == cudaSuccess &&
            cudaMalloc((void**)&p->t_hi, elems * 2) == cudaSuccess && cudaMalloc((void**)&p->t_lo, elems * 2) == cudaSuccess;
  if (ok) {
    cudaMemset(p->a_hi, 0, elems * 2); cudaMemset(p->a_lo, 0, elems * 2);
    cudaMemset(p->t_hi, 0, elems * 2); cudaMemset(p->t_lo, 0, elems * 2);
    ok = make_map_x(&p->tmA_hi, p->a_hi, (size_t)2 * p->CS * 128, p->KQ, 128) == 0 &&
         make_map_x(&p->tmA_lo, p->a_lo, (size_t)2 * p->CS * 128, p->KQ, 128) == 0 &&
         make_map_x(&p->tmT_hi, p->t_hi, (size_t)2 * p->CS * p->KQ, 128, 128) == 0 &&
         make_map_x(&p->tmT_lo, p->t_lo, (size_t)2 * p->CS * p->KQ, 128, 128) == 0;
  }
  if (ok) ok = p->tmem ? (tcx_attrs(lstm_tcx_fwd<true>, tcx_fwd_smem(p)) == cudaSuccess && tcx_attrs(lstm_tcx_bwd<true>, tcx_bwd_smem(p)) == cudaSuccess)
                       : (tcx_attrs(lstm_tcx_fwd<false>, tcx_fwd_smem(p)) == cudaSuccess && tcx_attrs(lstm_tcx_bwd<false>, tcx_bwd_smem(p)) == cudaSuccess);
  if (ok) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    int nf = 0, nb = 0;
    tcx_launch_cfg(p, cfg, at, 16, tcx_fwd_smem(p), nullptr);
    cudaError_t e = p->tmem ? cudaOccupancyMaxActiveClusters(&nf, lstm_tcx_fwd<true>, &cfg) : cudaOccupancyMaxActiveClusters(&nf, lstm_tcx_fwd<false>, &cfg);
    tcx_launch_cfg(p, cfg, at, 16, tcx_bwd_smem(p), nullptr);
    if (e == cudaSuccess) e = p->tmem ? cudaOccupancyMaxActiveClusters(&nb, lstm_tcx_bwd<true>, &cfg) : cudaOccupancyMaxActiveClusters(&nb, lstm_tcx_bwd<false>, &cfg);
    if (e != cudaSuccess || nf < 1 || nb < 1) ok = false;
    p->max_clusters = std::min(nf, nb);
    if (getenv("CLSTM_B200_TC_DBG")) fprintf(stderr, "lstm_tcx: nhidden %d -> clusters of %d CTAs, KQ %d, %s weights, resident clusters fwd %d bwd %d, smem fwd %zu bwd %zu\n",
                                             no, p->CS, p->KQ, p->tmem ? "TMEM" : "smem", nf, nb, tcx_fwd_smem(p), tcx_bwd_smem(p));
  }
  if (!ok) { cudaGetLastError(); lstm_tcx_destroy(p); return nullptr; }
  return p;
}

// 0: launched, -1: not applicable, > 0: CUDA error
int lstm_tcx_forward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  if (ensure_split_x(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  TcxArgs x{};
  tcx_fill(p, x, ln.B, a.d0, a.ndir, a.hstride, a.hoff);
  const int ncl = std::min(x.ngroups * a.ndir, p->max_clusters);
  if (ncl < 1) return -1;
  for (int d = 0; d < 2; d++) { x.XP[d] = a.XP[d]; x.G[d] = a.G[d]; x.C[d] = a.C[d]; x.Hprev[d] = a.Hprev[d]; }
  x.H = a.H; x.w_hi = p->a_hi; x.w_lo = p->a_lo;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  tcx_launch_cfg(p, cfg, at, ncl, tcx_fwd_smem(p), st);
  cudaError_t e = p->tmem ? cudaLaunchKernelEx(&cfg, lstm_tcx_fwd<true>, p->tmA_hi, p->tmA_lo, ln, x)
                          : cudaLaunchKernelEx(&cfg, lstm_tcx_fwd<false>, p->tmA_hi, p->tmA_lo, ln, x);
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tcx_fwd launch (%d clusters of %d): %s", ncl, p->CS, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

int lstm_tcx_backward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  if (ensure_split_x(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  TcxArgs x{};
  tcx_fill(p, x, ln.B, a.d0, a.ndir, a.hstride, a.hoff);
  const int ncl = std::min(x.ngroups * a.ndir, p->max_clusters);
  if (ncl < 1) return -1;
  for (int d = 0; d < 2; d++) { x.G[d] = const_cast<float*>(a.G[d]); x.C[d] = const_cast<float*>(a.C[d]); x.DG[d] = a.DG[d]; }
  x.dH = a.dH; x.w_hi = p->t_hi; x.w_lo = p->t_lo;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  tcx_launch_cfg(p, cfg, at, ncl, tcx_bwd_smem(p), st);
  cudaError_t e = p->tmem ? cudaLaunchKernelEx(&cfg, lstm_tcx_bwd<true>, p->tmT_hi, p->tmT_lo, ln, x)
                          : cudaLaunchKernelEx(&cfg, lstm_tcx_bwd<false>, p->tmT_hi, p->tmT_lo, ln, x);
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tcx_bwd launch (%d clusters of %d): %s", ncl, p->CS, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

}  // namespace cb200
