// lstm_tcx.cu -- the recurrence of the NPLSTM for mid-size nets (nhidden <= 256) as a CLUSTER-RESIDENT tensor-core kernel:
// a thread-block cluster owns a group of 16 text lines of one direction for the whole sequence, the recurrent matrix is
// split over the CTAs' shared memories, the per-step products run on tcgen05 (lines on the UMMA N dimension, accumulators
// in TMEM) and h travels between the CTAs through DISTRIBUTED SHARED MEMORY (st.async + mbarrier complete_tx) -- no L2 round
// trip inside a step, which is what bounds the lock-step kernels of lstm_tc.cu at ~5 us per step.
//
// Reference semantics (paths relative to /root/reference), identical to lstm.cu / lstm_tc.cu:
//   GenericNPLSTM<SIG,TANH,TANH>::forward   clstm.cc:600-621 (loop body :612-620), forward_lin1 clstm_compute.cc:286
//   GenericNPLSTM::backward                 clstm.cc:622-653 (loop body :629-650), backward_lin1 clstm_compute.cc:296
//
// Geometry.  CS CTAs per cluster, each owns UPC = nhidden / CS hidden units (<= 32) = 4 UPC gate rows (gate-interleaved,
// row 4j+g) as ONE UMMA M tile of 128 rows (rows beyond 4 UPC are zero).  The K dimension is laid out in "slots":
// k' = 32 c + j addresses unit j of CTA c (slots UPC..31 of a CTA are zero padding), KQ = 32 CS slots in all, so the 8 units a
// warp of CTA c produces form exactly one 16-byte chunk of the K-major operand row.
//   forward :  pre[128 rows x 16 lines] = R_slice[128 x KQ] (A, resident, TMA-loaded once) * h_{s-1}[16 lines x KQ]^T (B)
//   backward:  part[k' tile of 128 x 16 lines] = Rt_slice[KQ x 128 rows] (A, resident) * delta[16 lines x 128 rows]^T (B, local)
// Operands are fp16 hi/lo pairs (tc_common.cuh::split_f16); B stacks the hi rows (0..15) over the lo rows (16..31), so one
// MMA with N = 32 gives A_hi*[B_hi ; B_lo] and a second one with N = 16 adds A_lo*B_hi: two tcgen05.mma per 16 k.
// TMEM lane = gate row: the four gates of a unit sit in four adjacent lanes; a 4x4 register transpose inside the quad
// (4 shuffles per 4 lines) hands lane g the four gates of lines 4i+g, whose cell state it keeps in registers.
// Exchange: every warp packs its 8 units x 16 lines into 16-byte chunks (32 shuffles) and st.async's them into the B buffer
// of every CTA of the cluster; the stores complete bytes on the destination's mbarrier, the MMA warp of each CTA waits on
// its own barrier only.  Two B buffers; safe by data flow (a CTA sends step s only after it received all of step s-1).
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

constexpr int kXL = 16;            // lines per cluster
constexpr int kXThreads = 160;     // warps 0..3: epilogue (TMEM lane = gate row), warp 4: loader / MMA issuer
constexpr int kXSlots = 32;        // unit slots per CTA in the K layout
constexpr float kXScaleH = 16.f, kXScaleR = 16.f, kXScaleD = 256.f;

struct TcxArgs {
  int no, no4, UPC, CS, KQ, nkc;   // hidden units, gate rows, units per CTA, CTAs per cluster, K slots, 64-wide K chunks
  int ngroups, d0, ndir, hstride, hoff[2];
  const float* XP[2];
  float* G[2];
  float* C[2];
  float* Hprev[2];
  float* H;
  // backward
  const float* dH;
  float* DG[2];
};

__device__ __forceinline__ unsigned mapa_u32(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// 16-byte store into CTA `rank`'s shared memory that also completes 16 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(unsigned local_addr, unsigned local_bar, unsigned rank, unsigned a, unsigned b, unsigned c,
                                            unsigned d) {
  const unsigned raddr = mapa_u32(local_addr, rank), rbar = mapa_u32(local_bar, rank);
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(a),
               "r"(b), "r"(c), "r"(d), "r"(rbar)
               : "memory");
}
// wait on a barrier whose bytes arrive from other CTAs of the cluster
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity) {
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
// 4x4 transpose inside a quad of lanes: in a[m] = value of THIS lane's gate for line 4i+m; out b[k] = gate k of line 4i+g
__device__ __forceinline__ void quad_transpose(const float* a, float* b, int g) {
  const bool hi = (g & 2) != 0, lo = (g & 1) != 0;
  const float r0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2);
  const float r1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
  const float own0 = hi ? a[2] : a[0], own1 = hi ? a[3] : a[1];        // my gate, lines 2hi, 2hi+1
  const float ra = __shfl_xor_sync(0xffffffffu, lo ? own0 : own1, 1);   // lane g^1's gate for my line
  const float rb = __shfl_xor_sync(0xffffffffu, lo ? r0 : r1, 1);       // lane g^3's gate for my line
  const float mine = lo ? own1 : own0, par = lo ? r1 : r0;              // gates g and g^2 of line 4i+g
  // b[k] = value of gate k: k^g = 0 mine, 1 ra, 2 par, 3 rb  (selects with static k: no local-memory indexing)
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const bool xl = ((k & 1) != 0) != lo, xh = ((k & 2) != 0) != hi;
    b[k] = xh ? (xl ? rb : par) : (xl ? ra : mine);
  }
}

// ================================================================================================ forward
__global__ void __launch_bounds__(kXThreads, 1)
lstm_tcx_fwd(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, Lines ln, TcxArgs p) {
  extern __shared__ __align__(1024) unsigned char xs[];
  __shared__ __align__(8) unsigned long long bars[4];          // hbar[2] (h tiles), accbar, abar (weights)
  __shared__ unsigned tmem_base_s;
  __shared__ int lineT[kXL], lineOff[kXL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned CS = cluster_nctarank(), c = cluster_ctarank();
  const int cluster_id = blockIdx.x / (int)CS, nclusters = gridDim.x / (int)CS;
  const int cl_per_dir = nclusters / p.ndir;                   // clusters are bound to one direction (one weight slice)
  const int q = cluster_id / cl_per_dir, d = p.d0 + q;
  const unsigned smem0 = (smem_u32(xs) + 1023u) & ~1023u;
  const unsigned a_hi0 = smem0, a_lo0 = smem0 + (unsigned)p.nkc * 16384u;       // A: chunk kc at +kc*16384 ([128 rows][128 B])
  const unsigned b0 = a_lo0 + (unsigned)p.nkc * 16384u;                        // B buffer b at b0 + b*(nkc*4096): chunk kc [32 rows][128 B]
  const unsigned bbytes = (unsigned)p.nkc * 4096u;
  const unsigned bar0 = smem_u32(&bars[0]);
  const unsigned hbar0 = bar0, hbar1 = bar0 + 8, accbar = bar0 + 16, abar = bar0 + 24;
  const unsigned hbytes = CS * 4u * kXL * 2u * 16u;           // bytes a CTA receives per step: CS x 4 warps x 16 lines x 2 planes x 16 B

  if (tid == 0) {
    mbar_init(hbar0, 1); mbar_init(hbar1, 1); mbar_init(accbar, 1); mbar_init(abar, 1);
    mbar_init_fence();
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
  }
  // the B buffers must never hold NaN patterns (slots of padding units are never written)
  for (unsigned i = tid; i < 2 * bbytes / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(b0 + 16 * i), "r"(0u) : "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (elect_one()) {       // this CTA's weight slice, once: rows [(d*CS + c)*128, +128) of the slot-ordered copy
      mbar_expect_tx(abar, (unsigned)p.nkc * 2u * 16384u);
      for (int kc = 0; kc < p.nkc; kc++) {
        tma_load_2d(a_hi0 + kc * 16384, &tmA_hi, kc * 64, (d * (int)CS + (int)c) * 128, abar);
        tma_load_2d(a_lo0 + kc * 16384, &tmA_lo, kc * 64, (d * (int)CS + (int)c) * 128, abar);
      }
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                                           // barriers of every CTA initialised before any remote store
  const unsigned tmem_d = tmem_base_s;

  const int groups_per_cluster_stride = cl_per_dir;
  const int my_first_group = cluster_id - q * cl_per_dir;
  unsigned hph0 = 0, hph1 = 0, accph = 0;

  for (int group = my_first_group; group < p.ngroups; group += groups_per_cluster_stride) {
    const int l0 = group * kXL;
    if (tid < kXL) {
      const int li = (l0 + tid < ln.B) ? ln.order[l0 + tid] : -1;
      lineT[tid] = (li >= 0) ? ln.T[li] : 0;
      lineOff[tid] = (li >= 0) ? ln.off[li] : 0;
    }
    __syncthreads();
    const int Tg = lineT[0];

    if (warp == 4) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      const unsigned idesc32 = make_idesc_f16(128, 32), idesc16 = make_idesc_f16(128, 16);
      const unsigned long long dbase = make_desc(0);
      auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
      mbar_wait(abar, 0);
      for (int s = 1; s < Tg; s++) {
        const unsigned b = (unsigned)(s - 1) & 1u;              // h_{s-1} sits in buffer (s-1)&1
        const unsigned hb = b ? hbar1 : hbar0;
        if (elect_one()) mbar_expect_tx(hb, hbytes);
        __syncwarp();
        mbar_wait_cluster(hb, b ? hph1 : hph0);
        if (b) hph1 ^= 1; else hph0 ^= 1;
        fence_proxy_async_smem();                               // remote generic-proxy stores -> tensor-core reads
        tc_fence_after();
        if (elect_one()) {
          for (int kc = 0; kc < p.nkc; kc++) {
            const unsigned long long ah = desc_of(a_hi0 + kc * 16384), al = desc_of(a_lo0 + kc * 16384);
            const unsigned long long bh = desc_of(b0 + b * bbytes + kc * 4096);
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              mma_f16(tmem_d, ah + 2 * ks, bh + 2 * ks, idesc32, (kc > 0 || ks > 0) ? 1u : 0u);   // R_hi [h_hi ; h_lo]
              mma_f16(tmem_d, al + 2 * ks, bh + 2 * ks, idesc16, 1u);                             // + R_lo h_hi
            }
          }
          mma_commit(accbar);
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------------------------------------------------ epilogue warps
      const int j = tid >> 2, g = tid & 3;                      // unit slot of this CTA, gate
      const bool real = j < p.UPC;
      const int unit = (int)c * p.UPC + j;                      // hidden unit
      const int grow = 4 * unit + g;                            // gate row
      const int no = p.no, no4 = p.no4;
      const float* __restrict__ XPd = p.XP[d];
      float* __restrict__ Gd = p.G[d];
      float* __restrict__ Cd = p.C[d];
      float* __restrict__ Hpd = p.Hprev[d];
      float* __restrict__ Hd = p.H + p.hoff[d];
      const unsigned taddr = tmem_d + ((unsigned)(32 * warp) << 16);
      constexpr float inv_scale = 1.0f / (kXScaleH * kXScaleR);
      float cst[4] = {0.f, 0.f, 0.f, 0.f};                      // cell state of unit j for lines 4i+g
      // where this lane's 16-byte chunk goes in a B buffer: line = lane % 16, plane = lane / 16 (0 hi, 1 lo)
      const int xl = lane & 15, plane = lane >> 4;
      const unsigned xrow = (unsigned)(xl + 16 * plane);
      const unsigned xoff = (unsigned)((int)c >> 1) * 4096u + xrow * 128u + ((((unsigned)(4 * ((int)c & 1) + warp)) ^ (xrow & 7u)) << 4);
      for (int s = 0; s < Tg; s++) {
        float xp[kXL];
#pragma unroll
        for (int l = 0; l < kXL; l++) {
          const int Tl = lineT[l];
          xp[l] = (real && s < Tl) ? XPd[((size_t)lineOff[l] + (d ? Tl - 1 - s : s)) * no4 + grow] : 0.f;
        }
        float act[kXL];
        if (s > 0) {
          mbar_wait(accbar, accph);
          accph ^= 1;
          tc_fence_after();
          float acc[32];
          tmem_ld<32>(taddr, acc);
#pragma unroll
          for (int l = 0; l < kXL; l++) act[l] = fmaf(acc[l] + acc[kXL + l], inv_scale, xp[l]);
        } else {
#pragma unroll
          for (int l = 0; l < kXL; l++) act[l] = xp[l];
        }
#pragma unroll
        for (int l = 0; l < kXL; l++) act[l] = (g == 3) ? tanh_fast(act[l]) : sigmoid_fast(act[l]);   // forward_full1 clstm.cc:614-617
        float hh[4];
        unsigned hp[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float gt[4];
          quad_transpose(act + 4 * i, gt, g);                   // gi, gf, go, ci of line 4i+g
          const bool on = real && s < lineT[4 * i + g];
          if (on) {
            cst[i] = fmaf(gt[1], cst[i], gt[3] * gt[0]);        // forward_statemem clstm_compute.cc:504-508
            hh[i] = tanh_fast(cst[i]) * gt[2];                  // forward_nonlingate :530-537
          } else hh[i] = 0.f;
          unsigned short h16, l16;
          split_f16(hh[i] * kXScaleH, h16, l16);
          hp[i] = pack_h2(h16, l16);
        }
        if (s + 1 < Tg) {
          // this warp's 8 units x 16 lines as 16-byte chunks: lane (line xl, plane) collects unit k's value from lane 4k + xl%4
          unsigned v[8];
#pragma unroll
          for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const unsigned t = __shfl_sync(0xffffffffu, hp[i], 4 * k + (xl & 3));
              if ((xl >> 2) == i) v[k] = t;
            }
          }
          unsigned w4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const unsigned x0 = plane ? (v[2 * e] >> 16) : (v[2 * e] & 0xffffu);
            const unsigned x1 = plane ? (v[2 * e + 1] >> 16) : (v[2 * e + 1] & 0xffffu);
            w4[e] = x0 | (x1 << 16);
          }
          tc_fence_before();                                    // (the accumulator has been read: tcgen05.wait::ld above)
          const unsigned dstb = b0 + ((unsigned)s & 1u) * bbytes + xoff;
          const unsigned hb = (s & 1) ? hbar1 : hbar0;
          for (unsigned r = 0; r < CS; r++) st_async_v4(dstb, hb, r, w4[0], w4[1], w4[2], w4[3]);
        }
        // ---- stash for the backward pass and the dense products
#pragma unroll
        for (int l = 0; l < kXL; l++) {
          const int Tl = lineT[l];
          if (real && s < Tl) Gd[((size_t)lineOff[l] + (d ? Tl - 1 - s : s)) * no4 + grow] = act[l];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int l = 4 * i + g, Tl = lineT[l];
          if (real && s < Tl) {
            const size_t col = (size_t)lineOff[l] + (d ? Tl - 1 - s : s);
            Cd[col * no + unit] = cst[i];
            Hd[col * p.hstride + unit] = hh[i];
            if (s + 1 < Tl) Hpd[(col + (d ? -1 : 1)) * (size_t)no + unit] = hh[i];
            if (s == 0) Hpd[col * no + unit] = 0.f;
          }
        }
      }
    }
    __syncthreads();
    cluster_sync_all();        // nobody may send the next group's h_0 into a buffer a slower CTA is still multiplying
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(32) : "memory");
  }
}


// ================================================================================================ backward
// smem: A (Rt slice) [mt][kc] chunks of [128 k' rows][64 r'] hi, then lo | B = delta tile [2 kc][32 rows][128 B] | reduce
// buffers [2][CS src][32 slots][16 lines] fp32
__global__ void __launch_bounds__(kXThreads, 1)
lstm_tcx_bwd(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, Lines ln, TcxArgs p) {
  extern __shared__ __align__(1024) unsigned char xs[];
  __shared__ __align__(8) unsigned long long bars[5];          // pbar[2] (partial sums), accbar, abar (weights), bbar (delta tile)
  __shared__ unsigned tmem_base_s;
  __shared__ int lineT[kXL], lineOff[kXL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned CS = cluster_nctarank(), c = cluster_ctarank();
  const int cluster_id = blockIdx.x / (int)CS, nclusters = gridDim.x / (int)CS;
  const int cl_per_dir = nclusters / p.ndir;
  const int q = cluster_id / cl_per_dir, d = p.d0 + q;
  const int nmt = p.KQ / 128 > 0 ? p.KQ / 128 : 1;             // M tiles of outputs (k' slots)
  const int mrows = p.KQ < 128 ? p.KQ : 128;                   // (KQ = 64: one partial tile)
  const unsigned smem0 = (smem_u32(xs) + 1023u) & ~1023u;
  const unsigned a_hi0 = smem0, a_lo0 = smem0 + (unsigned)nmt * 2u * 16384u;     // chunk (mt, kc) at +(mt*2+kc)*16384
  const unsigned b0 = a_lo0 + (unsigned)nmt * 2u * 16384u;                      // delta tile: chunk kc at +kc*4096
  const unsigned r0 = b0 + 8192u;                                               // reduce buffers
  const unsigned rbytes = CS * 32u * kXL * 4u;
  const unsigned bar0 = smem_u32(&bars[0]);
  const unsigned pbar0 = bar0, pbar1 = bar0 + 8, accbar = bar0 + 16, abar = bar0 + 24, bbar = bar0 + 32;

  if (tid == 0) {
    mbar_init(pbar0, 1); mbar_init(pbar1, 1); mbar_init(accbar, 1); mbar_init(abar, 1); mbar_init(bbar, 128);
    mbar_init_fence();
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
  }
  for (unsigned i = tid; i < 8192 / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(b0 + 16 * i), "r"(0u) : "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (elect_one()) {       // Rt slice of this CTA: rows = output slots k', columns = this CTA's 128 gate rows
      mbar_expect_tx(abar, (unsigned)nmt * 2u * 2u * 16384u);
      for (int mt = 0; mt < nmt; mt++)
        for (int kc = 0; kc < 2; kc++) {
          const int row = (d * (int)CS + (int)c) * p.KQ + mt * 128;
          tma_load_2d(a_hi0 + (mt * 2 + kc) * 16384, &tmA_hi, kc * 64, row, abar);
          tma_load_2d(a_lo0 + (mt * 2 + kc) * 16384, &tmA_lo, kc * 64, row, abar);
        }
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const unsigned tmem_d = tmem_base_s;
  const int my_first_group = cluster_id - q * cl_per_dir;
  unsigned pph0 = 0, pph1 = 0, accph = 0, bph = 0;
  const unsigned pbytes = CS * 32u * kXL * 4u;                  // bytes a CTA receives per step (= one reduce buffer)

  for (int group = my_first_group; group < p.ngroups; group += cl_per_dir) {
    const int l0 = group * kXL;
    if (tid < kXL) {
      const int li = (l0 + tid < ln.B) ? ln.order[l0 + tid] : -1;
      lineT[tid] = (li >= 0) ? ln.T[li] : 0;
      lineOff[tid] = (li >= 0) ? ln.off[li] : 0;
    }
    __syncthreads();
    const int Tg = lineT[0];

    if (warp == 4) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      const unsigned idesc32 = make_idesc_f16(128, 32), idesc16 = make_idesc_f16(128, 16);
      const unsigned long long dbase = make_desc(0);
      auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
      mbar_wait(abar, 0);
      for (int fs = Tg - 1; fs >= 1; fs--) {
        mbar_wait(bbar, bph);                                   // the deltas of this step are in shared memory
        bph ^= 1;
        tc_fence_after();
        if (elect_one()) {
          for (int mt = 0; mt < nmt; mt++)
            for (int kc = 0; kc < 2; kc++) {
              const unsigned long long ah = desc_of(a_hi0 + (mt * 2 + kc) * 16384), al = desc_of(a_lo0 + (mt * 2 + kc) * 16384);
              const unsigned long long bh = desc_of(b0 + kc * 4096);
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                mma_f16(tmem_d + 32 * mt, ah + 2 * ks, bh + 2 * ks, idesc32, (kc > 0 || ks > 0) ? 1u : 0u);   // Rt_hi [d_hi ; d_lo]
                mma_f16(tmem_d + 32 * mt, al + 2 * ks, bh + 2 * ks, idesc16, 1u);                             // + Rt_lo d_hi
              }
            }
          mma_commit(accbar);
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------------------------------------------------ epilogue warps
      const int j = tid >> 2, g = tid & 3;
      const bool real = j < p.UPC;
      const int unit = (int)c * p.UPC + j;
      const int grow = 4 * unit + g;
      const int no = p.no, no4 = p.no4;
      const float* __restrict__ Gd = p.G[d];
      const float* __restrict__ Cd = p.C[d];
      const float* __restrict__ dHd = p.dH + p.hoff[d];
      float* __restrict__ DGd = p.DG[d];
      const unsigned taddr = tmem_d + ((unsigned)(32 * warp) << 16);
      constexpr float inv_scale = 1.0f / (kXScaleD * kXScaleR);
      float dcc[4] = {0.f, 0.f, 0.f, 0.f};                      // carried cell derivative of unit j for lines 4i+g
      // where this thread's deltas (gate row r' = tid, line l, plane) go in the delta tile
      const unsigned boff = (unsigned)(tid >> 6) * 4096u + 2u * (unsigned)(tid & 7);
      const unsigned bchunk = (unsigned)((tid & 63) >> 3);
      for (int it = 0; it < Tg; it++) {
        const int fs = Tg - 1 - it;
        // ---- operands that do not depend on the exchange
        float gact[kXL];
#pragma unroll
        for (int l = 0; l < kXL; l++) {
          const int Tl = lineT[l];
          gact[l] = (real && fs < Tl) ? Gd[((size_t)lineOff[l] + (d ? Tl - 1 - fs : fs)) * no4 + grow] : 0.f;
        }
        float cc[4], cp[4], dh[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int l = 4 * i + g, Tl = lineT[l];
          cc[i] = cp[i] = dh[i] = 0.f;
          if (real && fs < Tl) {
            const size_t col = (size_t)lineOff[l] + (d ? Tl - 1 - fs : fs);
            cc[i] = Cd[col * no + unit];
            if (fs > 0) cp[i] = Cd[(col + (d ? 1 : -1)) * (size_t)no + unit];
            dh[i] = dHd[col * p.hstride + unit];
          }
        }
        // ---- recurrent part of dh: the CS partial products of the previous step, fixed order
        if (it > 0) {
          const unsigned b = (unsigned)(it - 1) & 1u;
          const unsigned pb = b ? pbar1 : pbar0;
          if (tid == 0) mbar_expect_tx(pb, pbytes);
          mbar_wait_cluster(pb, b ? pph1 : pph0);
          if (b) pph1 ^= 1; else pph0 ^= 1;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int l = 4 * i + g;
            if (real && fs < lineT[l] - 1) {
              float r = 0.f;
              for (unsigned sc = 0; sc < CS; sc++) {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(r0 + b * rbytes + ((sc * 32u + (unsigned)j) * kXL + (unsigned)l) * 4u) : "memory");
                r += x;
              }
              dh[i] += r;
            }
          }
        }
        // ---- pointwise (backward_nonlingate / statemem / nonlin0, clstm_compute.cc:539-547, 509-515, 231-267)
        float dl[kXL];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float gt[4], dq[4];
          quad_transpose(gact + 4 * i, gt, g);                  // gi, gf, go, ci of line 4i+g
          const bool on = real && fs < lineT[4 * i + g];
          if (on) {
            const float th = tanh_fast(cc[i]);
            const float dgo = th * dh[i];
            const float dc = fmaf(1.f - th * th, gt[2] * dh[i], dcc[i]);
            float dgf = 0.f, carry = 0.f;
            if (fs > 0) { dgf = dc * cp[i]; carry = dc * gt[1]; }
            dcc[i] = carry;
            dq[0] = gt[0] * (1.f - gt[0]) * (dc * gt[3]);
            dq[1] = gt[1] * (1.f - gt[1]) * dgf;
            dq[2] = gt[2] * (1.f - gt[2]) * dgo;
            dq[3] = (1.f - gt[3] * gt[3]) * (dc * gt[0]);
          } else { dq[0] = dq[1] = dq[2] = dq[3] = 0.f; }
          quad_transpose(dq, dl + 4 * i, g);                    // back: this lane's gate for lines 4i..4i+3
        }
#pragma unroll
        for (int l = 0; l < kXL; l++) {
          const int Tl = lineT[l];
          if (real && fs < Tl) DGd[((size_t)lineOff[l] + (d ? Tl - 1 - fs : fs)) * no4 + grow] = dl[l];
        }
        if (fs == 0) break;
        // ---- delta tile (B operand): row = line (+16 for the lo plane), k = this thread's gate row
#pragma unroll
        for (int l = 0; l < kXL; l++) {
          unsigned short h16, l16;
          split_f16(dl[l] * kXScaleD, h16, l16);
          const unsigned rh = (unsigned)l, rl = (unsigned)l + 16u;
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(b0 + boff + rh * 128u + ((bchunk ^ (rh & 7u)) << 4)), "h"(h16) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(b0 + boff + rl * 128u + ((bchunk ^ (rl & 7u)) << 4)), "h"(l16) : "memory");
        }
        fence_proxy_async_smem();
        mbar_arrive(bbar);
        // ---- partial products of this CTA's gate rows for ALL output slots: send each slot's 16 lines to its owner
        mbar_wait(accbar, accph);
        accph ^= 1;
        tc_fence_after();
        for (int mt = 0; mt < nmt; mt++) {
          float acc[32];
          tmem_ld<32>(taddr + 32 * mt, acc);
          const unsigned owner = (unsigned)(4 * mt + warp);     // slots 128 mt + 32 warp .. +31 belong to this CTA of the cluster
          if (owner < CS && (mt * 128 + tid) < mrows * nmt) {
            const unsigned dst = r0 + ((unsigned)it & 1u) * rbytes + ((c * 32u + (unsigned)lane) * kXL) * 4u;
            const unsigned pb = (it & 1) ? pbar1 : pbar0;
#pragma unroll
            for (int e = 0; e < 4; e++)
              st_async_v4(dst + 16 * e, pb, owner, __float_as_uint((acc[4 * e] + acc[16 + 4 * e]) * inv_scale),
                          __float_as_uint((acc[4 * e + 1] + acc[16 + 4 * e + 1]) * inv_scale),
                          __float_as_uint((acc[4 * e + 2] + acc[16 + 4 * e + 2]) * inv_scale),
                          __float_as_uint((acc[4 * e + 3] + acc[16 + 4 * e + 3]) * inv_scale));
          }
        }
        tc_fence_before();
      }
    }
    __syncthreads();
    cluster_sync_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(64) : "memory");
  }
}

// ================================================================================================ weight layouts
// R [4no x no] fp32 -> slot-ordered fp16 hi/lo copy for the forward kernel: row (c*128 + 4j+g) = gate row 4(c*UPC+j)+g of CTA c,
// column k' = 32 c' + j' = unit c'*UPC + j'.  Padding stays zero from the allocation.
// Backward copy: for CTA c a [KQ output slots][128 gate rows of c] matrix, element (k', 4j+g) = the same weight.
__global__ void lstm_tcx_split_kernel(const float* __restrict__ R, int no, int UPC, int KQ, __half* __restrict__ a_hi,
                                      __half* __restrict__ a_lo, __half* __restrict__ t_hi, __half* __restrict__ t_lo) {
  const size_t total = (size_t)4 * no * no;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / no), k = (int)(i % no);
    const int unit = r >> 2, g = r & 3;
    const int c = unit / UPC, j = unit % UPC, c2 = k / UPC, j2 = k % UPC;
    unsigned short hi, lo;
    split_f16(R[i] * kXScaleR, hi, lo);
    const size_t o = (size_t)(c * 128 + 4 * j + g) * KQ + (32 * c2 + j2);
    reinterpret_cast<unsigned short*>(a_hi)[o] = hi;
    reinterpret_cast<unsigned short*>(a_lo)[o] = lo;
    const size_t ot = ((size_t)c * KQ + (32 * c2 + j2)) * 128 + (4 * j + g);
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
  int no = 0, num_sms = 148, UPC = 0, CS = 0, KQ = 0, nkc = 0;
  __half *a_hi = nullptr, *a_lo = nullptr;       // forward weight slices [2 dirs][CS][128][KQ]
  __half *t_hi = nullptr, *t_lo = nullptr;       // backward weight slices [2 dirs][CS][KQ][128]
  bool stale[2] = {true, true};
  CUtensorMap tmA_hi, tmA_lo, tmT_hi, tmT_lo;
  int max_clusters = 0;
  char err[256] = {0};
};

bool lstm_tcx_supported(int no) {
  if (no < 32 || no > 256 || no % 4 != 0) return false;
  for (int cs = 3; cs <= 8; cs++)          // (at least 3 CTAs: 128 output slots, one full partial-sum tile in the backward kernel)
    if (no % cs == 0 && no / cs <= kXSlots) return true;
  return false;
}
void lstm_tcx_destroy(LstmTcxPlan* p) {
  if (!p) return;
  cudaFree(p->a_hi); cudaFree(p->a_lo); cudaFree(p->t_hi); cudaFree(p->t_lo);
  delete p;
}
const char* lstm_tcx_error(const LstmTcxPlan* p) { return p ? p->err : "no plan"; }
void lstm_tcx_mark_stale(LstmTcxPlan* p) { if (p) p->stale[0] = p->stale[1] = true; }

namespace {
size_t tcx_fwd_smem(int nkc) { return (size_t)nkc * 2 * 16384 + (size_t)2 * nkc * 4096 + 1024; }
size_t tcx_bwd_smem(int KQ, int CS) {
  const int nmt = KQ / 128 > 0 ? KQ / 128 : 1;
  return (size_t)nmt * 2 * 2 * 16384 + 8192 + (size_t)2 * CS * 32 * kXL * 4 + 1024;
}
int ensure_split_x(LstmTcxPlan* p, cudaStream_t st, const float* const R[2], int d0, int ndir) {
  for (int d = d0; d < d0 + ndir; d++) {
    if (!p->stale[d]) continue;
    const size_t off = (size_t)d * p->CS * 128 * p->KQ;
    const size_t total = (size_t)4 * p->no * p->no;
    const int nb = (int)std::min<size_t>((total + 255) / 256, (size_t)p->num_sms * 8);
    lstm_tcx_split_kernel<<<nb, 256, 0, st>>>(R[d], p->no, p->UPC, p->KQ, p->a_hi + off, p->a_lo + off, p->t_hi + off, p->t_lo + off);
    p->stale[d] = false;
  }
  return (int)cudaGetLastError();
}
}

LstmTcxPlan* lstm_tcx_create(int no, int num_sms) {
  if (!lstm_tcx_supported(no) || load_encode_x() != 0) return nullptr;
  auto* p = new LstmTcxPlan;
  p->no = no; p->num_sms = num_sms;
  for (int cs = 3; cs <= 8; cs++)
    if (no % cs == 0 && no / cs <= kXSlots) { p->CS = cs; break; }       // the smallest cluster that holds the units
  p->UPC = no / p->CS; p->KQ = ((kXSlots * p->CS + 63) / 64) * 64; p->nkc = p->KQ / 64;
  const size_t elems = (size_t)2 * p->CS * 128 * p->KQ;
  bool ok = cudaMalloc((void**)&p->a_hi, elems * 2) == cudaSuccess && cudaMalloc((void**)&p->a_lo, elems * 2) == cudaSuccess &&
            cudaMalloc((void**)&p->t_hi, elems * 2) == cudaSuccess && cudaMalloc((void**)&p->t_lo, elems * 2) == cudaSuccess;
  if (ok) {
    cudaMemset(p->a_hi, 0, elems * 2); cudaMemset(p->a_lo, 0, elems * 2);
    cudaMemset(p->t_hi, 0, elems * 2); cudaMemset(p->t_lo, 0, elems * 2);
    const int trows = p->KQ < 128 ? p->KQ : 128;
    ok = make_map_x(&p->tmA_hi, p->a_hi, (size_t)2 * p->CS * 128, p->KQ, 128) == 0 &&
         make_map_x(&p->tmA_lo, p->a_lo, (size_t)2 * p->CS * 128, p->KQ, 128) == 0 &&
         make_map_x(&p->tmT_hi, p->t_hi, (size_t)2 * p->CS * p->KQ, 128, trows) == 0 &&
         make_map_x(&p->tmT_lo, p->t_lo, (size_t)2 * p->CS * p->KQ, 128, trows) == 0;
  }
  if (ok) ok = cudaFuncSetAttribute(lstm_tcx_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcx_fwd_smem(p->nkc)) == cudaSuccess &&
               cudaFuncSetAttribute(lstm_tcx_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcx_bwd_smem(p->KQ, p->CS)) == cudaSuccess;
  if (ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p->CS * 32); cfg.blockDim = dim3(kXThreads); cfg.dynamicSmemBytes = tcx_fwd_smem(p->nkc);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, lstm_tcx_fwd, &cfg) != cudaSuccess || nc < 1) ok = false;
    p->max_clusters = nc;
  }
  if (!ok) { cudaGetLastError(); lstm_tcx_destroy(p); return nullptr; }
  return p;
}

// 0: launched, -1: not applicable, > 0: CUDA error
int lstm_tcx_forward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  if (ensure_split_x(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  const int ngroups = (ln.B + kXL - 1) / kXL;
  int cl_per_dir = std::min(ngroups, p->max_clusters / a.ndir);
  if (cl_per_dir < 1) return -1;
  TcxArgs x{};
  x.no = p->no; x.no4 = 4 * p->no; x.UPC = p->UPC; x.CS = p->CS; x.KQ = p->KQ; x.nkc = p->nkc;
  x.ngroups = ngroups; x.d0 = a.d0; x.ndir = a.ndir; x.hstride = a.hstride; x.hoff[0] = a.hoff[0]; x.hoff[1] = a.hoff[1];
  for (int d = 0; d < 2; d++) { x.XP[d] = a.XP[d]; x.G[d] = a.G[d]; x.C[d] = a.C[d]; x.Hprev[d] = a.Hprev[d]; }
  x.H = a.H;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl_per_dir * a.ndir * p->CS); cfg.blockDim = dim3(kXThreads); cfg.dynamicSmemBytes = tcx_fwd_smem(p->nkc);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tcx_fwd, p->tmA_hi, p->tmA_lo, ln, x);
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tcx_fwd launch (%d clusters of %d): %s", cl_per_dir * a.ndir, p->CS, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

int lstm_tcx_backward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  if (ensure_split_x(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  const int ngroups = (ln.B + kXL - 1) / kXL;
  const int cl_per_dir = std::min(ngroups, p->max_clusters / a.ndir);
  if (cl_per_dir < 1) return -1;
  TcxArgs x{};
  x.no = p->no; x.no4 = 4 * p->no; x.UPC = p->UPC; x.CS = p->CS; x.KQ = p->KQ; x.nkc = p->nkc;
  x.ngroups = ngroups; x.d0 = a.d0; x.ndir = a.ndir; x.hstride = a.hstride; x.hoff[0] = a.hoff[0]; x.hoff[1] = a.hoff[1];
  for (int d = 0; d < 2; d++) { x.G[d] = const_cast<float*>(a.G[d]); x.C[d] = const_cast<float*>(a.C[d]); x.DG[d] = a.DG[d]; }
  x.dH = a.dH;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl_per_dir * a.ndir * p->CS); cfg.blockDim = dim3(kXThreads); cfg.dynamicSmemBytes = tcx_bwd_smem(p->KQ, p->CS);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tcx_bwd, p->tmT_hi, p->tmT_lo, ln, x);
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tcx_bwd launch (%d clusters of %d): %s", cl_per_dir * a.ndir, p->CS, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

}  // namespace cb200
