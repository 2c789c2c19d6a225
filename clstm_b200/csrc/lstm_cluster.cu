// lstm_cluster.cu -- register-resident recurrence for hidden sizes whose recurrent matrix does not fit one SM:
// a thread-block CLUSTER of CS CTAs per (line, direction).  CTA c owns UC = NO/CS hidden units, i.e. the 4*UC gate
// rows of those units times all NO columns = 40 000 weights = 100 per thread x 400 threads, exactly the register
// budget of the single-CTA kernel (lstm.cu).  NO = 200 -> CS = 4, NO = 400 -> CS = 16 (non-portable cluster size).
//
// forward : every CTA keeps a full copy of h (NO floats, double buffered) in its shared memory.  Per step a group of
//           LU = NO/25 lanes per unit splits K into slices of 25, reduce-scatters the four gate sums, applies the
//           gates / cell update, and lane i of the group stores the new h of the unit into CTA i's copy through
//           distributed shared memory (st.shared::cluster).  One cluster barrier per step.
// backward: each CTA multiplies ITS rows of R (own units) with its own deltas for ALL NO outputs -- no delta exchange --
//           and scatters the partial sums to the CTA that owns the output unit (DSMEM); the owner adds the CS partials.
//           One CTA barrier + one cluster barrier per step.
// Same math, gate-interleaved row order, fast gate functions and staging scheme as lstm.cu (see there for the
// reference citations).
#include <cstdlib>

#include "kernels.h"

namespace cb200 {
namespace {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float x, float y) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ void ffma2(u64& acc, u64 a, u64 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
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
__device__ __forceinline__ ulonglong2 lds_v2u64(unsigned addr) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// store into the shared memory of CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ void st_dsmem_f32(unsigned local_addr, unsigned rank, float v) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- mbarrier + st.async exchange: the store and the signal are one operation, nothing waits on a release fence
__device__ __forceinline__ void mbar_init_(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm_(unsigned bar, unsigned bytes) {   // this CTA's single arrival + the bytes to expect
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();       // a lost transfer must surface as an error, never as a hang
  }
}
// 4-byte store into CTA `rank`'s shared memory that also completes 4 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_f32(unsigned local_addr, unsigned local_bar, unsigned rank, float v) {
  unsigned raddr, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(__float_as_uint(v)), "r"(rbar)
               : "memory");
}

// split form: global stores issued between arrive and wait are not covered by this step's release, so the barrier
// does not have to wait for them to drain (they have a whole step until the next arrive)
__device__ __forceinline__ void cluster_arrive_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank_() {
  unsigned r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cp_async4(unsigned saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kStage = 4;

template <int NO> struct ClCfg {
  static constexpr int SL = 25, NPF = 12;              // k-slice per lane: 12 packed pairs + 1 tail
  static constexpr int SSTR = 28;                      // slice stride in smem (16-byte aligned, conflict-free)
  static constexpr int LU = NO / SL;                   // lanes per unit (8 or 16)
  static constexpr int CS = (NO * NO) / 10000;         // CTAs per cluster (4 or 16)
  static constexpr int UC = NO / CS;                   // units per CTA (50 or 25)
  static constexpr int THREADS = UC * LU;              // 400
  static constexpr int TPAD = (THREADS + 31) & ~31;    // 416
  static constexpr int LB = (LU == 8) ? 3 : 4;         // log2(LU)
  static_assert(NO % SL == 0 && (LU == 8 || LU == 16) && UC * CS == NO && THREADS == 400, "unsupported size");
};

// reduce-scatter four partials over the top two bits of the lane-in-group index, all-reduce over the remaining bits:
// afterwards every lane holds the group total of gate q = 2*bit(LB-1) + bit(LB-2).
template <int LB>
__device__ __forceinline__ float group_reduce4(float p0, float p1, float p2, float p3, bool hi, bool lo) {
  constexpr int X1 = 1 << (LB - 1), X2 = 1 << (LB - 2);
  float k0 = hi ? p2 : p0, k1 = hi ? p3 : p1;
  const float s0 = hi ? p0 : p2, s1 = hi ? p1 : p3;
  k0 += __shfl_xor_sync(0xffffffffu, s0, X1);
  k1 += __shfl_xor_sync(0xffffffffu, s1, X1);
  float r = lo ? k1 : k0;
  const float s2 = lo ? k0 : k1;
  r += __shfl_xor_sync(0xffffffffu, s2, X2);
#pragma unroll
  for (int x = X2 >> 1; x > 0; x >>= 1) r += __shfl_xor_sync(0xffffffffu, r, x);
  return r;
}

// ----------------------------------------------------------------------------------------------------- forward
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_fwd_cluster(Lines ln, LstmFwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, LU = Cfg::LU, CS = Cfg::CS, UC = Cfg::UC, LB = Cfg::LB;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD;
  __shared__ __align__(16) float h_s[2][LU * SSTR];
  __shared__ float xp_s[kStage][TPAD];
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int b = ln.order[blockIdx.x / CS], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - LU + (tid0 % LU);   // padding lanes clone the last unit group
  const int ul = tid / LU, lg = tid % LU;                          // unit inside the CTA, lane inside the unit group
  const int unit = (int)crank * UC + ul;                           // global hidden unit
  const bool hi = (lg >> (LB - 1)) & 1, lo = (lg >> (LB - 2)) & 1;
  const int q = 2 * (int)hi + (int)lo;                             // gate this lane ends up with
  const bool lead = (lg & ((1 << (LB - 2)) - 1)) == 0;             // one lane per (unit, gate)
  const int row = 4 * unit + q;
  const float* __restrict__ XPb = d ? a.XP[1] : a.XP[0];
  float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  float* __restrict__ Hpb = d ? a.Hprev[1] : a.Hprev[0];
  float* __restrict__ Hb = a.H + a.hoff[d];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* Rd = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float* Rr = Rd + (size_t)(4 * unit + g) * NO + lg * SL;
#pragma unroll
      for (int p = 0; p < NPF; p++) w[g][p] = pack2(Rr[2 * p], Rr[2 * p + 1]);
      wt[g] = Rr[SL - 1];
    }
  }
  for (int k = tid0; k < 2 * LU * SSTR; k += blockDim.x) (&h_s[0][0])[k] = 0.f;

  const unsigned hs_base = (unsigned)__cvta_generic_to_shared(&h_s[0][0]);
  constexpr unsigned BUFB = LU * SSTR * 4;
  unsigned rd_addr = hs_base + lg * (SSTR * 4);
  unsigned wr_addr = hs_base + BUFB + ((unit / SL) * SSTR + (unit % SL)) * 4;   // h[unit] in buffer 1 (any CTA)
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(&xp_s[0][tid0]);

  const int dt = d ? -1 : 1;
  unsigned ncol = off + (d ? T - 1 : 0);
  // output stream of the lead lanes: q=0 -> H, q=1 -> C, q=2 -> Hprev (h of the previous step), q=3 -> nothing extra
  float* __restrict__ obase = (q == 0) ? Hb : (q == 1) ? Cb : Hpb;
  const unsigned ostride = (q == 0) ? (unsigned)a.hstride : NO;

#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T) cp_async4(xs_addr + u * (TPAD * 4), XPb + (size_t)(ncol + u * dt) * ROWS + row);
    cp_async_commit();
  }
  cluster_sync_();        // every CTA's h buffers are zeroed before anybody writes into them remotely

  float c = 0.f, hprev = 0.f;
  const float sc = (q == 3) ? -2.f * kLog2e : -kLog2e;
  int tog = (int)BUFB;
  for (int s = 0; s < T; s++) {
    {
      const int sn = s + kStage - 1;
      if (sn < T) cp_async4(xs_addr + (sn & (kStage - 1)) * (TPAD * 4), XPb + (size_t)(ncol + (kStage - 1) * dt) * ROWS + row);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float htail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 h2 = lds_v2u64(rd_addr + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], h2.x); ffma2(acc1, w[1][2 * i], h2.x);
        ffma2(acc2, w[2][2 * i], h2.x); ffma2(acc3, w[3][2 * i], h2.x);
        ffma2(acc0, w[0][2 * i + 1], h2.y); ffma2(acc1, w[1][2 * i + 1], h2.y);
        ffma2(acc2, w[2][2 * i + 1], h2.y); ffma2(acc3, w[3][2 * i + 1], h2.y);
      } else {
        float dummy; unpack2(h2.x, htail, dummy);
      }
    }
    const float xp = lds_f32(xs_addr + (s & (kStage - 1)) * (TPAD * 4));
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], htail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], htail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], htail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], htail, l0 + l1);
    const float pre = group_reduce4<LB>(p0, p1, p2, p3, hi, lo) + xp;
    const float sg = rcp_approx(1.0f + ex2_approx(sc * pre));
    const float act = (q == 3) ? fmaf(2.f, sg, -1.f) : sg;
    const float gi = __shfl_sync(0xffffffffu, act, 0 << (LB - 2), LU);
    const float gf = __shfl_sync(0xffffffffu, act, 1 << (LB - 2), LU);
    const float go = __shfl_sync(0xffffffffu, act, 2 << (LB - 2), LU);
    const float ci = __shfl_sync(0xffffffffu, act, 3 << (LB - 2), LU);
    c = fmaf(gf, c, ci * gi);
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    const float hh = th * go;
    if (lg < CS) st_dsmem_f32(wr_addr, (unsigned)lg, hh);           // lane i of the group feeds CTA i's copy of h
    cluster_arrive_();
    if (lead) Gb[ncol * ROWS + row] = act;                          // stash stores: after the arrive (see cluster_arrive_)
    if (lead && q < 3) obase[(size_t)ncol * ostride + unit] = (q == 1) ? c : (q == 2) ? hprev : hh;
    hprev = hh;
    ncol += dt;
    cluster_wait_();
    rd_addr += tog; wr_addr -= tog; tog = -tog;
  }
}

// ----------------------------------------------------------------------------------------------------- backward
// thread = (k-group kgp of 4 outputs, row slice rs of 25 own rows): 100 weights R[own row][k].  KG = NO/4 groups,
// RSL = 4*UC/25 row slices per group (8 for NO=200, 4 for NO=400): KG*RSL = 400 threads.
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_bwd_cluster(Lines ln, LstmBwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, CS = Cfg::CS, UC = Cfg::UC;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD, OWN = 4 * UC;     // own delta rows per CTA (200 or 100)
  constexpr int RSL = OWN / SL;                                    // 8 or 4 row slices
  constexpr int RB = (RSL == 8) ? 3 : 2;
  extern __shared__ __align__(16) float bsm[];
  float* dg_s = bsm;                         // [2][RSL * SSTR]   own deltas, sliced
  float* part_s = bsm + 2 * RSL * SSTR;      // [2][CS][UC]       partial dh of my units from every CTA (double buffered)
  float* st_s = part_s + 2 * CS * UC;        // [kStage][TPAD][8] per-thread staging ring
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int b = ln.order[blockIdx.x / CS], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const unsigned st_addr0 = (unsigned)__cvta_generic_to_shared(st_s + (size_t)tid0 * 8);
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - RSL + (tid0 % RSL);   // padding lanes clone the last k-group
  // ---- matvec role: outputs kg4..kg4+3 (global unit indices), own-row slice rs
  const int kgp = tid / RSL, rs = tid % RSL;
  const int kg4 = kgp * 4;
  const bool hi = (rs >> (RB - 1)) & 1, lo = (rs >> (RB - 2)) & 1;
  const int kout = kg4 + 2 * (int)hi + (int)lo;                    // output unit this lane holds after the reduce
  // ---- pointwise role: thread -> (own unit pu, gate pg), 4*UC <= 400 threads take part
  const bool pw = tid0 < OWN;
  const int pu = pw ? tid0 >> 2 : 0, pg = tid0 & 3;
  const int punit = (int)crank * UC + pu;
  const float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  const float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  const float* __restrict__ dHb = a.dH + a.hoff[d];
  float* __restrict__ DGb = d ? a.DG[1] : a.DG[0];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* R = d ? a.R[1] : a.R[0];
    const int r0 = 4 * (int)crank * UC + rs * SL;                  // first own row of the slice (global row index)
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int p = 0; p < NPF; p++)
        w[kk][p] = pack2(R[(size_t)(r0 + 2 * p) * NO + kg4 + kk], R[(size_t)(r0 + 2 * p + 1) * NO + kg4 + kk]);
      wt[kk] = R[(size_t)(r0 + SL - 1) * NO + kg4 + kk];
    }
  }
  for (int i = tid0; i < 2 * RSL * SSTR + 2 * CS * UC; i += blockDim.x) bsm[i] = 0.f;

  const unsigned ds_base = (unsigned)__cvta_generic_to_shared(dg_s);
  constexpr unsigned DBUF = RSL * SSTR * 4;
  unsigned rd_addr = ds_base + rs * (SSTR * 4);
  const int prow = 4 * pu + pg;                                    // own delta row published by the pointwise thread
  unsigned wr_addr = ds_base + ((prow / SL) * SSTR + (prow % SL)) * 4;
  const unsigned ps_base = (unsigned)__cvta_generic_to_shared(part_s);
  constexpr unsigned PBUF = CS * UC * 4;
  // where this lane's reduced partial goes: CTA kout/UC, slot [crank][kout % UC]
  const unsigned dst_rank = (unsigned)(kout / UC);
  unsigned pdst = ps_base + ((int)crank * UC + (kout % UC)) * 4;   // buffer 0
  unsigned psrc = ps_base + pu * 4;                                // + c*UC*4 per source CTA, buffer 1 first (zeros)
  const bool send = (rs & ((1 << (RB - 2)) - 1)) == 0 || RB == 2;  // one lane per output after the all-reduce
  constexpr unsigned STG = TPAD * 32;

  const int dt = d ? 1 : -1;
  unsigned ncol = off + (d ? 0 : T - 1);
  auto stage = [&](int u, unsigned col) {
    const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
    cp_async16(sa, Gb + (size_t)col * ROWS + 4 * punit);
    cp_async4(sa + 16, Cb + (size_t)col * NO + punit);
    if (u + 1 < T) cp_async4(sa + 20, Cb + (size_t)(col + dt) * NO + punit);
    cp_async4(sa + 24, dHb + (size_t)col * a.hstride + punit);
  };
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T && pw) stage(u, ncol + u * dt);
    cp_async_commit();
  }
  cluster_sync_();

  float dcc = 0.f;
  const bool p_lo = (pg & 1) != 0, p_hi = (pg & 2) != 0;
  int dtog = (int)DBUF, ptog = (int)PBUF;
  unsigned pread = psrc + PBUF;                                    // read buffer 1 (zero) at the first step
  for (int u = 0; u < T; u++) {
    {
      const int un = u + kStage - 1;
      if (un < T && pw) stage(un, ncol + (kStage - 1) * dt);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const bool first = (u + 1 == T);
    float dl_keep = 0.f;
    if (pw) {
      const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
      const ulonglong2 gq = lds_v2u64(sa);
      const ulonglong2 cq = lds_v2u64(sa + 16);
      float gi, gf, go, ci, c, cprev, dhu, unused;
      unpack2(gq.x, gi, gf); unpack2(gq.y, go, ci);
      unpack2(cq.x, c, cprev); unpack2(cq.y, dhu, unused);
      if (first) cprev = 0.f;
      float dhrec = 0.f;
#pragma unroll
      for (int cc = 0; cc < CS; cc++) dhrec += lds_f32(pread + cc * (UC * 4));   // partials from every CTA, fixed order
      const float dh = dhu + dhrec;
      const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
      const float dc = fmaf(1.f - th * th, go * dh, dcc);
      dcc = first ? 0.f : dc * gf;
      const float y0 = p_lo ? gf : gi, y1 = p_lo ? ci : go;
      const float y = p_hi ? y1 : y0;
      const float B0 = p_lo ? cprev : ci, B1 = p_lo ? gi : dh;
      const float Bv = p_hi ? B1 : B0;
      const float Av = (pg == 2) ? th : dc;
      const float fp = (1.f - y) * ((pg == 3) ? (1.f + y) : y);
      const float dl = fp * (Av * Bv);
      sts_f32(wr_addr, dl);
      dl_keep = dl;
    }
    __syncthreads();
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float dtail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 d2 = lds_v2u64(rd_addr + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], d2.x); ffma2(acc1, w[1][2 * i], d2.x);
        ffma2(acc2, w[2][2 * i], d2.x); ffma2(acc3, w[3][2 * i], d2.x);
        ffma2(acc0, w[0][2 * i + 1], d2.y); ffma2(acc1, w[1][2 * i + 1], d2.y);
        ffma2(acc2, w[2][2 * i + 1], d2.y); ffma2(acc3, w[3][2 * i + 1], d2.y);
      } else {
        float dummy; unpack2(d2.x, dtail, dummy);
      }
    }
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], dtail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], dtail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], dtail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], dtail, l0 + l1);
    const float part = group_reduce4<RB>(p0, p1, p2, p3, hi, lo);
    if (send && tid0 < Cfg::THREADS) st_dsmem_f32(pdst, dst_rank, part);       // my rows' share of dh_prev[kout]
    cluster_arrive_();
    if (pw) DGb[ncol * ROWS + 4 * punit + pg] = dl_keep;                        // after the arrive (see cluster_arrive_)
    ncol += dt;
    cluster_wait_();
    rd_addr += dtog; wr_addr += dtog; dtog = -dtog;
    pread = psrc + ((ptog > 0) ? 0u : PBUF);                                    // the buffer that was just filled
    pdst += ptog; ptog = -ptog;
  }
}

// Backward, exchange through mbarriers (see lstm_fwd_cluster3): the partial sums travel with st.async and complete the
// owner's barrier; only the pointwise threads wait for them, the CTA barrier that follows orders everybody else.
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_bwd_cluster3(Lines ln, LstmBwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, CS = Cfg::CS, UC = Cfg::UC;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD, OWN = 4 * UC;     // own delta rows per CTA (200 or 100)
  constexpr int RSL = OWN / SL;                                    // 8 or 4 row slices
  constexpr int RB = (RSL == 8) ? 3 : 2;
  extern __shared__ __align__(16) float bsm[];
  float* dg_s = bsm;                         // [2][RSL * SSTR]   own deltas, sliced
  float* part_s = bsm + 2 * RSL * SSTR;      // [2][CS][UC]       partial dh of my units from every CTA (double buffered)
  float* st_s = part_s + 2 * CS * UC;        // [kStage][TPAD][8] per-thread staging ring
  __shared__ __align__(8) unsigned long long pbar[2];   // one mbarrier per partial-sum buffer
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int b = ln.order[blockIdx.x / CS], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const unsigned st_addr0 = (unsigned)__cvta_generic_to_shared(st_s + (size_t)tid0 * 8);
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - RSL + (tid0 % RSL);   // padding lanes clone the last k-group
  // ---- matvec role: outputs kg4..kg4+3 (global unit indices), own-row slice rs
  const int kgp = tid / RSL, rs = tid % RSL;
  const int kg4 = kgp * 4;
  const bool hi = (rs >> (RB - 1)) & 1, lo = (rs >> (RB - 2)) & 1;
  const int kout = kg4 + 2 * (int)hi + (int)lo;                    // output unit this lane holds after the reduce
  // ---- pointwise role: thread -> (own unit pu, gate pg), 4*UC <= 400 threads take part
  const bool pw = tid0 < OWN;
  const int pu = pw ? tid0 >> 2 : 0, pg = tid0 & 3;
  const int punit = (int)crank * UC + pu;
  const float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  const float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  const float* __restrict__ dHb = a.dH + a.hoff[d];
  float* __restrict__ DGb = d ? a.DG[1] : a.DG[0];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* R = d ? a.R[1] : a.R[0];
    const int r0 = 4 * (int)crank * UC + rs * SL;                  // first own row of the slice (global row index)
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int p = 0; p < NPF; p++)
        w[kk][p] = pack2(R[(size_t)(r0 + 2 * p) * NO + kg4 + kk], R[(size_t)(r0 + 2 * p + 1) * NO + kg4 + kk]);
      wt[kk] = R[(size_t)(r0 + SL - 1) * NO + kg4 + kk];
    }
  }
  for (int i = tid0; i < 2 * RSL * SSTR + 2 * CS * UC; i += blockDim.x) bsm[i] = 0.f;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&pbar[0]);
  if (tid0 == 0) {
    mbar_init_(bar0, 1);
    mbar_init_(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const unsigned ds_base = (unsigned)__cvta_generic_to_shared(dg_s);
  constexpr unsigned DBUF = RSL * SSTR * 4;
  unsigned rd_addr = ds_base + rs * (SSTR * 4);
  const int prow = 4 * pu + pg;                                    // own delta row published by the pointwise thread
  unsigned wr_addr = ds_base + ((prow / SL) * SSTR + (prow % SL)) * 4;
  const unsigned ps_base = (unsigned)__cvta_generic_to_shared(part_s);
  constexpr unsigned PBUF = CS * UC * 4;
  // where this lane's reduced partial goes: CTA kout/UC, slot [crank][kout % UC]
  const unsigned dst_rank = (unsigned)(kout / UC);
  unsigned pdst = ps_base + ((int)crank * UC + (kout % UC)) * 4;   // buffer 0
  unsigned psrc = ps_base + pu * 4;                                // + c*UC*4 per source CTA, buffer 1 first (zeros)
  const bool send = (rs & ((1 << (RB - 2)) - 1)) == 0 || RB == 2;  // one lane per output after the all-reduce
  constexpr unsigned STG = TPAD * 32;

  const int dt = d ? 1 : -1;
  unsigned ncol = off + (d ? 0 : T - 1);
  auto stage = [&](int u, unsigned col) {
    const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
    cp_async16(sa, Gb + (size_t)col * ROWS + 4 * punit);
    cp_async4(sa + 16, Cb + (size_t)col * NO + punit);
    if (u + 1 < T) cp_async4(sa + 20, Cb + (size_t)(col + dt) * NO + punit);
    cp_async4(sa + 24, dHb + (size_t)col * a.hstride + punit);
  };
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T && pw) stage(u, ncol + u * dt);
    cp_async_commit();
  }
  cluster_sync_();

  float dcc = 0.f;
  const bool p_lo = (pg & 1) != 0, p_hi = (pg & 2) != 0;
  int dtog = (int)DBUF, ptog = (int)PBUF;
  unsigned pread = psrc + PBUF;                                    // read buffer 1 (zero) at the first step
  for (int u = 0; u < T; u++) {
    {
      const int un = u + kStage - 1;
      if (un < T && pw) stage(un, ncol + (kStage - 1) * dt);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const bool first = (u + 1 == T);
    float dl_keep = 0.f;
    const unsigned wbar = bar0 + ((u & 1) << 3);                   // barrier of the buffer that receives this step's partials
    if (tid0 == 0 && !first) mbar_arm_(wbar, NO * 4);
    if (pw) {
      if (u > 0) mbar_wait_(bar0 + (((u - 1) & 1) << 3), ((unsigned)(u - 1) >> 1) & 1u);   // partials of step u-1 are complete
      const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
      const ulonglong2 gq = lds_v2u64(sa);
      const ulonglong2 cq = lds_v2u64(sa + 16);
      float gi, gf, go, ci, c, cprev, dhu, unused;
      unpack2(gq.x, gi, gf); unpack2(gq.y, go, ci);
      unpack2(cq.x, c, cprev); unpack2(cq.y, dhu, unused);
      if (first) cprev = 0.f;
      float dhrec = 0.f;
#pragma unroll
      for (int cc = 0; cc < CS; cc++) dhrec += lds_f32(pread + cc * (UC * 4));   // partials from every CTA, fixed order
      const float dh = dhu + dhrec;
      const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
      const float dc = fmaf(1.f - th * th, go * dh, dcc);
      dcc = first ? 0.f : dc * gf;
      const float y0 = p_lo ? gf : gi, y1 = p_lo ? ci : go;
      const float y = p_hi ? y1 : y0;
      const float B0 = p_lo ? cprev : ci, B1 = p_lo ? gi : dh;
      const float Bv = p_hi ? B1 : B0;
      const float Av = (pg == 2) ? th : dc;
      const float fp = (1.f - y) * ((pg == 3) ? (1.f + y) : y);
      const float dl = fp * (Av * Bv);
      sts_f32(wr_addr, dl);
      dl_keep = dl;
    }
    __syncthreads();
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float dtail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 d2 = lds_v2u64(rd_addr + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], d2.x); ffma2(acc1, w[1][2 * i], d2.x);
        ffma2(acc2, w[2][2 * i], d2.x); ffma2(acc3, w[3][2 * i], d2.x);
        ffma2(acc0, w[0][2 * i + 1], d2.y); ffma2(acc1, w[1][2 * i + 1], d2.y);
        ffma2(acc2, w[2][2 * i + 1], d2.y); ffma2(acc3, w[3][2 * i + 1], d2.y);
      } else {
        float dummy; unpack2(d2.x, dtail, dummy);
      }
    }
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], dtail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], dtail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], dtail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], dtail, l0 + l1);
    const float part = group_reduce4<RB>(p0, p1, p2, p3, hi, lo);
    if (!first && send && tid0 < Cfg::THREADS) st_async_f32(pdst, wbar, dst_rank, part);   // my rows' share of dh_prev[kout]
    if (pw) DGb[ncol * ROWS + 4 * punit + pg] = dl_keep;
    ncol += dt;
    rd_addr += dtog; wr_addr += dtog; dtog = -dtog;
    pread = psrc + ((ptog > 0) ? 0u : PBUF);                                    // the buffer that was just filled
    pdst += ptog; ptog = -ptog;
  }
  cluster_sync_();        // nobody leaves while a peer might still be using the cluster's shared memory windows
}


// ---------------------------------------------------------------------------------------------------------------------
// Forward, exchange through mbarriers: every CTA owns one mbarrier per h buffer.  At step s a CTA arms the barrier of the
// buffer that will receive h_s (one arrival + NO*4 expected bytes); the lanes that used to store h remotely now issue
// st.async, which delivers the value AND completes 4 bytes on the destination's barrier; step s+1 starts with a wait on
// the local barrier.  No cluster-wide rendezvous and no release fence in the loop: a CTA only waits for the data it
// reads.  Double buffering is safe by data flow (a CTA sends h_s only after it has received all of h_{s-1}, i.e. after
// every peer finished reading the buffer that h_s overwrites).
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_fwd_cluster3(Lines ln, LstmFwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, LU = Cfg::LU, CS = Cfg::CS, UC = Cfg::UC, LB = Cfg::LB;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD;
  __shared__ __align__(16) float h_s[2][LU * SSTR];
  __shared__ float xp_s[kStage][TPAD];
  __shared__ __align__(8) unsigned long long hbar[2];
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int b = ln.order[blockIdx.x / CS], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - LU + (tid0 % LU);
  const int ul = tid / LU, lg = tid % LU;
  const int unit = (int)crank * UC + ul;
  const bool hi = (lg >> (LB - 1)) & 1, lo = (lg >> (LB - 2)) & 1;
  const int q = 2 * (int)hi + (int)lo;
  const bool lead = (lg & ((1 << (LB - 2)) - 1)) == 0;
  const int row = 4 * unit + q;
  const float* __restrict__ XPb = d ? a.XP[1] : a.XP[0];
  float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  float* __restrict__ Hpb = d ? a.Hprev[1] : a.Hprev[0];
  float* __restrict__ Hb = a.H + a.hoff[d];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* Rd = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float* Rr = Rd + (size_t)(4 * unit + g) * NO + lg * SL;
#pragma unroll
      for (int p = 0; p < NPF; p++) w[g][p] = pack2(Rr[2 * p], Rr[2 * p + 1]);
      wt[g] = Rr[SL - 1];
    }
  }
  for (int k = tid0; k < 2 * LU * SSTR; k += blockDim.x) (&h_s[0][0])[k] = 0.f;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&hbar[0]);
  if (tid0 == 0) {
    mbar_init_(bar0, 1);
    mbar_init_(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const unsigned hs_base = (unsigned)__cvta_generic_to_shared(&h_s[0][0]);
  constexpr unsigned BUFB = LU * SSTR * 4;
  unsigned rd_addr = hs_base + lg * (SSTR * 4);
  unsigned wr_addr = hs_base + BUFB + ((unit / SL) * SSTR + (unit % SL)) * 4;
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(&xp_s[0][tid0]);
  const int dt = d ? -1 : 1;
  unsigned ncol = off + (d ? T - 1 : 0);
  float* __restrict__ obase = (q == 0) ? Hb : (q == 1) ? Cb : Hpb;
  const unsigned ostride = (q == 0) ? (unsigned)a.hstride : NO;
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T) cp_async4(xs_addr + u * (TPAD * 4), XPb + (size_t)(ncol + u * dt) * ROWS + row);
    cp_async_commit();
  }
  cluster_sync_();        // buffers zeroed and barriers initialised in every CTA before the first remote store

  float c = 0.f, hprev = 0.f;
  const float sc = (q == 3) ? -2.f * kLog2e : -kLog2e;
  int tog = (int)BUFB;
  for (int s = 0; s < T; s++) {
    {
      const int sn = s + kStage - 1;
      if (sn < T) cp_async4(xs_addr + (sn & (kStage - 1)) * (TPAD * 4), XPb + (size_t)(ncol + (kStage - 1) * dt) * ROWS + row);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const unsigned wbar = bar0 + (((s + 1) & 1) << 3);             // barrier of the buffer that receives h_s
    if (tid0 == 0 && s + 1 < T) mbar_arm_(wbar, NO * 4);
    if (s > 0) mbar_wait_(bar0 + ((s & 1) << 3), ((unsigned)(s - 1) >> 1) & 1u);   // h_{s-1} complete in this CTA
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float htail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 h2 = lds_v2u64(rd_addr + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], h2.x); ffma2(acc1, w[1][2 * i], h2.x);
        ffma2(acc2, w[2][2 * i], h2.x); ffma2(acc3, w[3][2 * i], h2.x);
        ffma2(acc0, w[0][2 * i + 1], h2.y); ffma2(acc1, w[1][2 * i + 1], h2.y);
        ffma2(acc2, w[2][2 * i + 1], h2.y); ffma2(acc3, w[3][2 * i + 1], h2.y);
      } else {
        float dummy; unpack2(h2.x, htail, dummy);
      }
    }
    const float xp = lds_f32(xs_addr + (s & (kStage - 1)) * (TPAD * 4));
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], htail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], htail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], htail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], htail, l0 + l1);
    const float pre = group_reduce4<LB>(p0, p1, p2, p3, hi, lo) + xp;
    const float sg = rcp_approx(1.0f + ex2_approx(sc * pre));
    const float act = (q == 3) ? fmaf(2.f, sg, -1.f) : sg;
    const float gi = __shfl_sync(0xffffffffu, act, 0 << (LB - 2), LU);
    const float gf = __shfl_sync(0xffffffffu, act, 1 << (LB - 2), LU);
    const float go = __shfl_sync(0xffffffffu, act, 2 << (LB - 2), LU);
    const float ci = __shfl_sync(0xffffffffu, act, 3 << (LB - 2), LU);
    c = fmaf(gf, c, ci * gi);
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    const float hh = th * go;
    if (s + 1 < T && lg < CS && tid0 < Cfg::THREADS) st_async_f32(wr_addr, wbar, (unsigned)lg, hh);
    if (lead) Gb[ncol * ROWS + row] = act;
    if (lead && q < 3) obase[(size_t)ncol * ostride + unit] = (q == 1) ? c : (q == 2) ? hprev : hh;
    hprev = hh;
    ncol += dt;
    rd_addr += tog; wr_addr -= tog; tog = -tog;
  }
  cluster_sync_();        // nobody leaves while a peer might still be using the cluster's shared memory windows
}

// =====================================================================================================================
// Two lines per cluster, software pipelined.  One recurrence step of a cluster is dominated by the exchange (DSMEM store
// latency + cluster barrier), not by its ~1100 cycles of arithmetic; with more (line, direction) chains than resident
// clusters the exchange of line A is hidden behind the arithmetic of line B and vice versa.  The weights in registers
// serve both lines.  The single hardware cluster barrier is used in strictly alternating phases:
//      compute A(s) | wait Y(s-1) | send h_A(s), arrive X(s) | compute B(s) | wait X(s) | send h_B(s), arrive Y(s) | ...
// Lines are paired in decreasing-length order (ln.order), so line A is never shorter than line B; once B has ended its
// slots are skipped (all CTAs of the cluster see the same lengths, so the barrier sequence stays uniform).
// =====================================================================================================================
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_fwd_cluster2(Lines ln, LstmFwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, LU = Cfg::LU, CS = Cfg::CS, UC = Cfg::UC, LB = Cfg::LB;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD;
  __shared__ __align__(16) float h_s[2][2][LU * SSTR];             // [line][buffer]
  __shared__ float xp_s[2][kStage][TPAD];                          // [line][stage]
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int pair = blockIdx.x / CS, d = a.d0 + blockIdx.y;
  const bool hasB = 2 * pair + 1 < ln.B;
  const int bA = ln.order[2 * pair], bB = ln.order[hasB ? 2 * pair + 1 : 2 * pair];
  const int TA = ln.T[bA], TB = hasB ? ln.T[bB] : 0;
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - LU + (tid0 % LU);
  const int ul = tid / LU, lg = tid % LU;
  const int unit = (int)crank * UC + ul;
  const bool hi = (lg >> (LB - 1)) & 1, lo = (lg >> (LB - 2)) & 1;
  const int q = 2 * (int)hi + (int)lo;
  const bool lead = (lg & ((1 << (LB - 2)) - 1)) == 0;
  const int row = 4 * unit + q;
  const float* __restrict__ XPb = d ? a.XP[1] : a.XP[0];
  float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  float* __restrict__ Hpb = d ? a.Hprev[1] : a.Hprev[0];
  float* __restrict__ Hb = a.H + a.hoff[d];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* Rd = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float* Rr = Rd + (size_t)(4 * unit + g) * NO + lg * SL;
#pragma unroll
      for (int p = 0; p < NPF; p++) w[g][p] = pack2(Rr[2 * p], Rr[2 * p + 1]);
      wt[g] = Rr[SL - 1];
    }
  }
  for (int k = tid0; k < 4 * LU * SSTR; k += blockDim.x) (&h_s[0][0][0])[k] = 0.f;

  const unsigned hs_base = (unsigned)__cvta_generic_to_shared(&h_s[0][0][0]);
  constexpr unsigned BUFB = LU * SSTR * 4, LINEB = 2 * BUFB;       // bytes per h buffer / per line
  unsigned rd_addr = hs_base + lg * (SSTR * 4);                    // line A, buffer 0 (line B: + LINEB)
  unsigned wr_addr = hs_base + BUFB + ((unit / SL) * SSTR + (unit % SL)) * 4;
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(&xp_s[0][0][tid0]);
  constexpr unsigned XLINE = kStage * TPAD * 4;

  const int dt = d ? -1 : 1;
  unsigned ncolA = ln.off[bA] + (d ? TA - 1 : 0), ncolB = ln.off[bB] + (d ? max(TB, 1) - 1 : 0);
  float* __restrict__ obase = (q == 0) ? Hb : (q == 1) ? Cb : Hpb;
  const unsigned ostride = (q == 0) ? (unsigned)a.hstride : NO;

#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < TA) cp_async4(xs_addr + u * (TPAD * 4), XPb + (size_t)(ncolA + u * dt) * ROWS + row);
    if (u < TB) cp_async4(xs_addr + XLINE + u * (TPAD * 4), XPb + (size_t)(ncolB + u * dt) * ROWS + row);
    cp_async_commit();
  }
  cluster_sync_();

  float cA = 0.f, cB = 0.f, hpA = 0.f, hpB = 0.f;
  const float sc = (q == 3) ? -2.f * kLog2e : -kLog2e;
  int tog = (int)BUFB;
  // one recurrence step of one line: returns h, leaves the gate activation in `act` and the new cell in `c`
  auto compute = [&](unsigned rd, unsigned xs, float& c, float& act) -> float {
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float htail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 h2 = lds_v2u64(rd + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], h2.x); ffma2(acc1, w[1][2 * i], h2.x);
        ffma2(acc2, w[2][2 * i], h2.x); ffma2(acc3, w[3][2 * i], h2.x);
        ffma2(acc0, w[0][2 * i + 1], h2.y); ffma2(acc1, w[1][2 * i + 1], h2.y);
        ffma2(acc2, w[2][2 * i + 1], h2.y); ffma2(acc3, w[3][2 * i + 1], h2.y);
      } else {
        float dummy; unpack2(h2.x, htail, dummy);
      }
    }
    const float xp = lds_f32(xs);
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], htail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], htail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], htail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], htail, l0 + l1);
    const float pre = group_reduce4<LB>(p0, p1, p2, p3, hi, lo) + xp;
    const float sg = rcp_approx(1.0f + ex2_approx(sc * pre));
    act = (q == 3) ? fmaf(2.f, sg, -1.f) : sg;
    const float gi = __shfl_sync(0xffffffffu, act, 0 << (LB - 2), LU);
    const float gf = __shfl_sync(0xffffffffu, act, 1 << (LB - 2), LU);
    const float go = __shfl_sync(0xffffffffu, act, 2 << (LB - 2), LU);
    const float ci = __shfl_sync(0xffffffffu, act, 3 << (LB - 2), LU);
    c = fmaf(gf, c, ci * gi);
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    return th * go;
  };
  for (int s = 0; s < TA; s++) {
    {
      const int sn = s + kStage - 1;
      const unsigned slot = (sn & (kStage - 1)) * (TPAD * 4);
      if (sn < TA) cp_async4(xs_addr + slot, XPb + (size_t)(ncolA + (kStage - 1) * dt) * ROWS + row);
      if (sn < TB) cp_async4(xs_addr + XLINE + slot, XPb + (size_t)(ncolB + (kStage - 1) * dt) * ROWS + row);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const unsigned xslot = (s & (kStage - 1)) * (TPAD * 4);
    const bool liveB = s < TB;
    // ---- line A
    float actA;
    const float hA = compute(rd_addr, xs_addr + xslot, cA, actA);
    if (s > 0) cluster_wait_();                                    // Y(s-1): everybody has line B's previous h
    if (lg < CS) st_dsmem_f32(wr_addr, (unsigned)lg, hA);
    cluster_arrive_();                                             // X(s)
    if (lead) Gb[ncolA * ROWS + row] = actA;
    if (lead && q < 3) obase[(size_t)ncolA * ostride + unit] = (q == 1) ? cA : (q == 2) ? hpA : hA;
    hpA = hA;
    ncolA += dt;
    // ---- line B (its arithmetic covers the exchange of line A)
    float actB = 0.f, hB = 0.f;
    if (liveB) hB = compute(rd_addr + LINEB, xs_addr + XLINE + xslot, cB, actB);
    cluster_wait_();                                               // X(s)
    if (liveB && lg < CS) st_dsmem_f32(wr_addr + LINEB, (unsigned)lg, hB);
    cluster_arrive_();                                             // Y(s)
    if (liveB) {
      if (lead) Gb[ncolB * ROWS + row] = actB;
      if (lead && q < 3) obase[(size_t)ncolB * ostride + unit] = (q == 1) ? cB : (q == 2) ? hpB : hB;
      hpB = hB;
      ncolB += dt;
    }
    rd_addr += tog; wr_addr -= tog; tog = -tog;
  }
  cluster_wait_();                                                 // the last Y
}

template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_bwd_cluster2(Lines ln, LstmBwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, CS = Cfg::CS, UC = Cfg::UC;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD, OWN = 4 * UC;
  constexpr int RSL = OWN / SL;
  constexpr int RB = (RSL == 8) ? 3 : 2;
  extern __shared__ __align__(16) float bsm[];
  float* dg_s = bsm;                         // [line][2][RSL * SSTR]
  float* part_s = bsm + 4 * RSL * SSTR;      // [line][2][CS][UC]
  float* st_s = part_s + 4 * CS * UC;        // [line][kStage][TPAD][8]
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int pair = blockIdx.x / CS, d = a.d0 + blockIdx.y;
  const bool hasB = 2 * pair + 1 < ln.B;
  const int bA = ln.order[2 * pair], bB = ln.order[hasB ? 2 * pair + 1 : 2 * pair];
  const int TA = ln.T[bA], TB = hasB ? ln.T[bB] : 0;
  const unsigned st_addr0 = (unsigned)__cvta_generic_to_shared(st_s + (size_t)tid0 * 8);
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - RSL + (tid0 % RSL);
  const int kgp = tid / RSL, rs = tid % RSL;
  const int kg4 = kgp * 4;
  const bool hi = (rs >> (RB - 1)) & 1, lo = (rs >> (RB - 2)) & 1;
  const int kout = kg4 + 2 * (int)hi + (int)lo;
  const bool pw = tid0 < OWN;
  const int pu = pw ? tid0 >> 2 : 0, pg = tid0 & 3;
  const int punit = (int)crank * UC + pu;
  const float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  const float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  const float* __restrict__ dHb = a.dH + a.hoff[d];
  float* __restrict__ DGb = d ? a.DG[1] : a.DG[0];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* R = d ? a.R[1] : a.R[0];
    const int r0 = 4 * (int)crank * UC + rs * SL;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int p = 0; p < NPF; p++)
        w[kk][p] = pack2(R[(size_t)(r0 + 2 * p) * NO + kg4 + kk], R[(size_t)(r0 + 2 * p + 1) * NO + kg4 + kk]);
      wt[kk] = R[(size_t)(r0 + SL - 1) * NO + kg4 + kk];
    }
  }
  for (int i = tid0; i < 4 * RSL * SSTR + 4 * CS * UC; i += blockDim.x) bsm[i] = 0.f;

  const unsigned ds_base = (unsigned)__cvta_generic_to_shared(dg_s);
  constexpr unsigned DBUF = RSL * SSTR * 4, DLINE = 2 * DBUF;
  unsigned rd_addr = ds_base + rs * (SSTR * 4);
  const int prow = 4 * pu + pg;
  unsigned wr_addr = ds_base + ((prow / SL) * SSTR + (prow % SL)) * 4;
  const unsigned ps_base = (unsigned)__cvta_generic_to_shared(part_s);
  constexpr unsigned PBUF = CS * UC * 4, PLINE = 2 * PBUF;
  const unsigned dst_rank = (unsigned)(kout / UC);
  unsigned pdst = ps_base + ((int)crank * UC + (kout % UC)) * 4;   // line A, buffer 0
  const unsigned psrc = ps_base + pu * 4;
  const bool send = (rs & ((1 << (RB - 2)) - 1)) == 0 || RB == 2;
  constexpr unsigned STG = TPAD * 32, SLINE = kStage * STG;

  const int dt = d ? 1 : -1;
  unsigned ncolA = ln.off[bA] + (d ? 0 : TA - 1), ncolB = ln.off[bB] + (d ? 0 : max(TB, 1) - 1);
  auto stage = [&](int u, unsigned col, int T, unsigned lineoff) {
    const unsigned sa = st_addr0 + lineoff + (u & (kStage - 1)) * STG;
    cp_async16(sa, Gb + (size_t)col * ROWS + 4 * punit);
    cp_async4(sa + 16, Cb + (size_t)col * NO + punit);
    if (u + 1 < T) cp_async4(sa + 20, Cb + (size_t)(col + dt) * NO + punit);
    cp_async4(sa + 24, dHb + (size_t)col * a.hstride + punit);
  };
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (pw && u < TA) stage(u, ncolA + u * dt, TA, 0);
    if (pw && u < TB) stage(u, ncolB + u * dt, TB, SLINE);
    cp_async_commit();
  }
  cluster_sync_();

  float dccA = 0.f, dccB = 0.f;
  const bool p_lo = (pg & 1) != 0, p_hi = (pg & 2) != 0;
  int dtog = (int)DBUF, ptog = (int)PBUF;
  unsigned pread = psrc + PBUF;                                    // line A: read buffer 1 (zeros) at the first step
  // pointwise part of one step of one line: publishes delta[prow] in shared memory, returns it
  auto pointwise = [&](unsigned sa, unsigned pr, unsigned wr, bool first, float& dcc) -> float {
    const ulonglong2 gq = lds_v2u64(sa);
    const ulonglong2 cq = lds_v2u64(sa + 16);
    float gi, gf, go, ci, c, cprev, dhu, unused;
    unpack2(gq.x, gi, gf); unpack2(gq.y, go, ci);
    unpack2(cq.x, c, cprev); unpack2(cq.y, dhu, unused);
    if (first) cprev = 0.f;
    float dhrec = 0.f;
#pragma unroll
    for (int cc = 0; cc < CS; cc++) dhrec += lds_f32(pr + cc * (UC * 4));
    const float dh = dhu + dhrec;
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    const float dc = fmaf(1.f - th * th, go * dh, dcc);
    dcc = first ? 0.f : dc * gf;
    const float y0 = p_lo ? gf : gi, y1 = p_lo ? ci : go;
    const float y = p_hi ? y1 : y0;
    const float B0 = p_lo ? cprev : ci, B1 = p_lo ? gi : dh;
    const float Bv = p_hi ? B1 : B0;
    const float Av = (pg == 2) ? th : dc;
    const float fp = (1.f - y) * ((pg == 3) ? (1.f + y) : y);
    const float dl = fp * (Av * Bv);
    sts_f32(wr, dl);
    return dl;
  };
  auto matvec = [&](unsigned rd) -> float {
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float dtail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 d2 = lds_v2u64(rd + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], d2.x); ffma2(acc1, w[1][2 * i], d2.x);
        ffma2(acc2, w[2][2 * i], d2.x); ffma2(acc3, w[3][2 * i], d2.x);
        ffma2(acc0, w[0][2 * i + 1], d2.y); ffma2(acc1, w[1][2 * i + 1], d2.y);
        ffma2(acc2, w[2][2 * i + 1], d2.y); ffma2(acc3, w[3][2 * i + 1], d2.y);
      } else {
        float dummy; unpack2(d2.x, dtail, dummy);
      }
    }
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], dtail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], dtail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], dtail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], dtail, l0 + l1);
    return group_reduce4<RB>(p0, p1, p2, p3, hi, lo);
  };
  for (int u = 0; u < TA; u++) {
    {
      const int un = u + kStage - 1;
      if (pw && un < TA) stage(un, ncolA + (kStage - 1) * dt, TA, 0);
      if (pw && un < TB) stage(un, ncolB + (kStage - 1) * dt, TB, SLINE);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const unsigned sslot = (u & (kStage - 1)) * STG;
    const bool liveB = u < TB;
    // ---- line A
    float dlA = 0.f;
    if (pw) dlA = pointwise(st_addr0 + sslot, pread, wr_addr, u + 1 == TA, dccA);
    __syncthreads();
    const float partA = matvec(rd_addr);
    if (u > 0) cluster_wait_();                                    // Y(u-1)
    if (send && tid0 < Cfg::THREADS) st_dsmem_f32(pdst, dst_rank, partA);
    cluster_arrive_();                                             // X(u)
    if (pw) DGb[ncolA * ROWS + 4 * punit + pg] = dlA;
    ncolA += dt;
    // ---- line B
    float dlB = 0.f, partB = 0.f;
    if (liveB && pw) dlB = pointwise(st_addr0 + SLINE + sslot, pread + PLINE, wr_addr + DLINE, u + 1 == TB, dccB);
    __syncthreads();
    if (liveB) partB = matvec(rd_addr + DLINE);
    cluster_wait_();                                               // X(u)
    if (liveB && send && tid0 < Cfg::THREADS) st_dsmem_f32(pdst + PLINE, dst_rank, partB);
    cluster_arrive_();                                             // Y(u)
    if (liveB) {
      if (pw) DGb[ncolB * ROWS + 4 * punit + pg] = dlB;
      ncolB += dt;
    }
    rd_addr += dtog; wr_addr += dtog; dtog = -dtog;
    pread = psrc + ((ptog > 0) ? 0u : PBUF);
    pdst += ptog; ptog = -ptog;
  }
  cluster_wait_();
}

// Two lines per cluster AND the mbarrier exchange: line A's h (partials) travel while line B is being computed, nobody
// waits at a rendezvous.  Same pairing rules as lstm_*_cluster2, same protocol per line as lstm_*_cluster3.
template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_fwd_cluster4(Lines ln, LstmFwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, LU = Cfg::LU, CS = Cfg::CS, UC = Cfg::UC, LB = Cfg::LB;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD;
  __shared__ __align__(16) float h_s[2][2][LU * SSTR];             // [line][buffer]
  __shared__ float xp_s[2][kStage][TPAD];                          // [line][stage]
  __shared__ __align__(8) unsigned long long hbar[2][2];           // [line][buffer]
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int pair = blockIdx.x / CS, d = a.d0 + blockIdx.y;
  const bool hasB = 2 * pair + 1 < ln.B;
  const int bA = ln.order[2 * pair], bB = ln.order[hasB ? 2 * pair + 1 : 2 * pair];
  const int TA = ln.T[bA], TB = hasB ? ln.T[bB] : 0;
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - LU + (tid0 % LU);
  const int ul = tid / LU, lg = tid % LU;
  const int unit = (int)crank * UC + ul;
  const bool hi = (lg >> (LB - 1)) & 1, lo = (lg >> (LB - 2)) & 1;
  const int q = 2 * (int)hi + (int)lo;
  const bool lead = (lg & ((1 << (LB - 2)) - 1)) == 0;
  const int row = 4 * unit + q;
  const float* __restrict__ XPb = d ? a.XP[1] : a.XP[0];
  float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  float* __restrict__ Hpb = d ? a.Hprev[1] : a.Hprev[0];
  float* __restrict__ Hb = a.H + a.hoff[d];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* Rd = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float* Rr = Rd + (size_t)(4 * unit + g) * NO + lg * SL;
#pragma unroll
      for (int p = 0; p < NPF; p++) w[g][p] = pack2(Rr[2 * p], Rr[2 * p + 1]);
      wt[g] = Rr[SL - 1];
    }
  }
  for (int k = tid0; k < 4 * LU * SSTR; k += blockDim.x) (&h_s[0][0][0])[k] = 0.f;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&hbar[0][0]);    // line A: bar0 + 8*buf, line B: bar0 + 16 + 8*buf
  if (tid0 == 0) {
#pragma unroll
    for (int k = 0; k < 4; k++) mbar_init_(bar0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const unsigned hs_base = (unsigned)__cvta_generic_to_shared(&h_s[0][0][0]);
  constexpr unsigned BUFB = LU * SSTR * 4, LINEB = 2 * BUFB;       // bytes per h buffer / per line
  unsigned rd_addr = hs_base + lg * (SSTR * 4);                    // line A, buffer 0 (line B: + LINEB)
  unsigned wr_addr = hs_base + BUFB + ((unit / SL) * SSTR + (unit % SL)) * 4;
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(&xp_s[0][0][tid0]);
  constexpr unsigned XLINE = kStage * TPAD * 4;

  const int dt = d ? -1 : 1;
  unsigned ncolA = ln.off[bA] + (d ? TA - 1 : 0), ncolB = ln.off[bB] + (d ? max(TB, 1) - 1 : 0);
  float* __restrict__ obase = (q == 0) ? Hb : (q == 1) ? Cb : Hpb;
  const unsigned ostride = (q == 0) ? (unsigned)a.hstride : NO;

#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < TA) cp_async4(xs_addr + u * (TPAD * 4), XPb + (size_t)(ncolA + u * dt) * ROWS + row);
    if (u < TB) cp_async4(xs_addr + XLINE + u * (TPAD * 4), XPb + (size_t)(ncolB + u * dt) * ROWS + row);
    cp_async_commit();
  }
  cluster_sync_();

  float cA = 0.f, cB = 0.f, hpA = 0.f, hpB = 0.f;
  const float sc = (q == 3) ? -2.f * kLog2e : -kLog2e;
  int tog = (int)BUFB;
  // one recurrence step of one line: returns h, leaves the gate activation in `act` and the new cell in `c`
  auto compute = [&](unsigned rd, unsigned xs, float& c, float& act) -> float {
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float htail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 h2 = lds_v2u64(rd + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], h2.x); ffma2(acc1, w[1][2 * i], h2.x);
        ffma2(acc2, w[2][2 * i], h2.x); ffma2(acc3, w[3][2 * i], h2.x);
        ffma2(acc0, w[0][2 * i + 1], h2.y); ffma2(acc1, w[1][2 * i + 1], h2.y);
        ffma2(acc2, w[2][2 * i + 1], h2.y); ffma2(acc3, w[3][2 * i + 1], h2.y);
      } else {
        float dummy; unpack2(h2.x, htail, dummy);
      }
    }
    const float xp = lds_f32(xs);
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], htail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], htail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], htail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], htail, l0 + l1);
    const float pre = group_reduce4<LB>(p0, p1, p2, p3, hi, lo) + xp;
    const float sg = rcp_approx(1.0f + ex2_approx(sc * pre));
    act = (q == 3) ? fmaf(2.f, sg, -1.f) : sg;
    const float gi = __shfl_sync(0xffffffffu, act, 0 << (LB - 2), LU);
    const float gf = __shfl_sync(0xffffffffu, act, 1 << (LB - 2), LU);
    const float go = __shfl_sync(0xffffffffu, act, 2 << (LB - 2), LU);
    const float ci = __shfl_sync(0xffffffffu, act, 3 << (LB - 2), LU);
    c = fmaf(gf, c, ci * gi);
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    return th * go;
  };
  for (int s = 0; s < TA; s++) {
    {
      const int sn = s + kStage - 1;
      const unsigned slot = (sn & (kStage - 1)) * (TPAD * 4);
      if (sn < TA) cp_async4(xs_addr + slot, XPb + (size_t)(ncolA + (kStage - 1) * dt) * ROWS + row);
      if (sn < TB) cp_async4(xs_addr + XLINE + slot, XPb + (size_t)(ncolB + (kStage - 1) * dt) * ROWS + row);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const unsigned xslot = (s & (kStage - 1)) * (TPAD * 4);
    const bool liveB = s < TB;
    const unsigned wsel = (unsigned)((s + 1) & 1) << 3, rsel = (unsigned)(s & 1) << 3;
    const unsigned par = ((unsigned)(s - 1) >> 1) & 1u;
    const bool sender = lg < CS && tid0 < Cfg::THREADS;
    if (tid0 == 0) {
      if (s + 1 < TA) mbar_arm_(bar0 + wsel, NO * 4);
      if (s + 1 < TB) mbar_arm_(bar0 + 16 + wsel, NO * 4);
    }
    // ---- line A (its h of the previous step travelled while line B was being computed)
    if (s > 0) mbar_wait_(bar0 + rsel, par);
    float actA;
    const float hA = compute(rd_addr, xs_addr + xslot, cA, actA);
    if (s + 1 < TA && sender) st_async_f32(wr_addr, bar0 + wsel, (unsigned)lg, hA);
    if (lead) Gb[ncolA * ROWS + row] = actA;
    if (lead && q < 3) obase[(size_t)ncolA * ostride + unit] = (q == 1) ? cA : (q == 2) ? hpA : hA;
    hpA = hA;
    ncolA += dt;
    // ---- line B (its arithmetic covers the exchange of line A)
    float actB = 0.f, hB = 0.f;
    if (liveB) {
      if (s > 0) mbar_wait_(bar0 + 16 + rsel, par);
      hB = compute(rd_addr + LINEB, xs_addr + XLINE + xslot, cB, actB);
      if (s + 1 < TB && sender) st_async_f32(wr_addr + LINEB, bar0 + 16 + wsel, (unsigned)lg, hB);
      if (lead) Gb[ncolB * ROWS + row] = actB;
      if (lead && q < 3) obase[(size_t)ncolB * ostride + unit] = (q == 1) ? cB : (q == 2) ? hpB : hB;
      hpB = hB;
      ncolB += dt;
    }
    rd_addr += tog; wr_addr -= tog; tog = -tog;
  }
  cluster_sync_();        // nobody leaves while a peer might still be using the cluster's shared memory windows
}

template <int NO>
__global__ void __launch_bounds__(ClCfg<NO>::TPAD, 1) lstm_bwd_cluster4(Lines ln, LstmBwdArgs a) {
  typedef ClCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NPF = Cfg::NPF, SSTR = Cfg::SSTR, CS = Cfg::CS, UC = Cfg::UC;
  constexpr int ROWS = 4 * NO, TPAD = Cfg::TPAD, OWN = 4 * UC;
  constexpr int RSL = OWN / SL;
  constexpr int RB = (RSL == 8) ? 3 : 2;
  extern __shared__ __align__(16) float bsm[];
  float* dg_s = bsm;                         // [line][2][RSL * SSTR]
  float* part_s = bsm + 4 * RSL * SSTR;      // [line][2][CS][UC]
  float* st_s = part_s + 4 * CS * UC;        // [line][kStage][TPAD][8]
  __shared__ __align__(8) unsigned long long pbar[2][2];           // [line][buffer]
  const int tid0 = threadIdx.x;
  const unsigned crank = cluster_rank_();
  const int pair = blockIdx.x / CS, d = a.d0 + blockIdx.y;
  const bool hasB = 2 * pair + 1 < ln.B;
  const int bA = ln.order[2 * pair], bB = ln.order[hasB ? 2 * pair + 1 : 2 * pair];
  const int TA = ln.T[bA], TB = hasB ? ln.T[bB] : 0;
  const unsigned st_addr0 = (unsigned)__cvta_generic_to_shared(st_s + (size_t)tid0 * 8);
  const int tid = (tid0 < Cfg::THREADS) ? tid0 : Cfg::THREADS - RSL + (tid0 % RSL);
  const int kgp = tid / RSL, rs = tid % RSL;
  const int kg4 = kgp * 4;
  const bool hi = (rs >> (RB - 1)) & 1, lo = (rs >> (RB - 2)) & 1;
  const int kout = kg4 + 2 * (int)hi + (int)lo;
  const bool pw = tid0 < OWN;
  const int pu = pw ? tid0 >> 2 : 0, pg = tid0 & 3;
  const int punit = (int)crank * UC + pu;
  const float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  const float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  const float* __restrict__ dHb = a.dH + a.hoff[d];
  float* __restrict__ DGb = d ? a.DG[1] : a.DG[0];

  u64 w[4][NPF];
  float wt[4];
  {
    const float* R = d ? a.R[1] : a.R[0];
    const int r0 = 4 * (int)crank * UC + rs * SL;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int p = 0; p < NPF; p++)
        w[kk][p] = pack2(R[(size_t)(r0 + 2 * p) * NO + kg4 + kk], R[(size_t)(r0 + 2 * p + 1) * NO + kg4 + kk]);
      wt[kk] = R[(size_t)(r0 + SL - 1) * NO + kg4 + kk];
    }
  }
  for (int i = tid0; i < 4 * RSL * SSTR + 4 * CS * UC; i += blockDim.x) bsm[i] = 0.f;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&pbar[0][0]);    // line A: bar0 + 8*buf, line B: bar0 + 16 + 8*buf
  if (tid0 == 0) {
#pragma unroll
    for (int k = 0; k < 4; k++) mbar_init_(bar0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const unsigned ds_base = (unsigned)__cvta_generic_to_shared(dg_s);
  constexpr unsigned DBUF = RSL * SSTR * 4, DLINE = 2 * DBUF;
  unsigned rd_addr = ds_base + rs * (SSTR * 4);
  const int prow = 4 * pu + pg;
  unsigned wr_addr = ds_base + ((prow / SL) * SSTR + (prow % SL)) * 4;
  const unsigned ps_base = (unsigned)__cvta_generic_to_shared(part_s);
  constexpr unsigned PBUF = CS * UC * 4, PLINE = 2 * PBUF;
  const unsigned dst_rank = (unsigned)(kout / UC);
  unsigned pdst = ps_base + ((int)crank * UC + (kout % UC)) * 4;   // line A, buffer 0
  const unsigned psrc = ps_base + pu * 4;
  const bool send = (rs & ((1 << (RB - 2)) - 1)) == 0 || RB == 2;
  constexpr unsigned STG = TPAD * 32, SLINE = kStage * STG;

  const int dt = d ? 1 : -1;
  unsigned ncolA = ln.off[bA] + (d ? 0 : TA - 1), ncolB = ln.off[bB] + (d ? 0 : max(TB, 1) - 1);
  auto stage = [&](int u, unsigned col, int T, unsigned lineoff) {
    const unsigned sa = st_addr0 + lineoff + (u & (kStage - 1)) * STG;
    cp_async16(sa, Gb + (size_t)col * ROWS + 4 * punit);
    cp_async4(sa + 16, Cb + (size_t)col * NO + punit);
    if (u + 1 < T) cp_async4(sa + 20, Cb + (size_t)(col + dt) * NO + punit);
    cp_async4(sa + 24, dHb + (size_t)col * a.hstride + punit);
  };
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (pw && u < TA) stage(u, ncolA + u * dt, TA, 0);
    if (pw && u < TB) stage(u, ncolB + u * dt, TB, SLINE);
    cp_async_commit();
  }
  cluster_sync_();

  float dccA = 0.f, dccB = 0.f;
  const bool p_lo = (pg & 1) != 0, p_hi = (pg & 2) != 0;
  int dtog = (int)DBUF, ptog = (int)PBUF;
  unsigned pread = psrc + PBUF;                                    // line A: read buffer 1 (zeros) at the first step
  // pointwise part of one step of one line: publishes delta[prow] in shared memory, returns it
  auto pointwise = [&](unsigned sa, unsigned pr, unsigned wr, bool first, float& dcc) -> float {
    const ulonglong2 gq = lds_v2u64(sa);
    const ulonglong2 cq = lds_v2u64(sa + 16);
    float gi, gf, go, ci, c, cprev, dhu, unused;
    unpack2(gq.x, gi, gf); unpack2(gq.y, go, ci);
    unpack2(cq.x, c, cprev); unpack2(cq.y, dhu, unused);
    if (first) cprev = 0.f;
    float dhrec = 0.f;
#pragma unroll
    for (int cc = 0; cc < CS; cc++) dhrec += lds_f32(pr + cc * (UC * 4));
    const float dh = dhu + dhrec;
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    const float dc = fmaf(1.f - th * th, go * dh, dcc);
    dcc = first ? 0.f : dc * gf;
    const float y0 = p_lo ? gf : gi, y1 = p_lo ? ci : go;
    const float y = p_hi ? y1 : y0;
    const float B0 = p_lo ? cprev : ci, B1 = p_lo ? gi : dh;
    const float Bv = p_hi ? B1 : B0;
    const float Av = (pg == 2) ? th : dc;
    const float fp = (1.f - y) * ((pg == 3) ? (1.f + y) : y);
    const float dl = fp * (Av * Bv);
    sts_f32(wr, dl);
    return dl;
  };
  auto matvec = [&](unsigned rd) -> float {
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float dtail = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const ulonglong2 d2 = lds_v2u64(rd + 16 * i);
      if (i < 6) {
        ffma2(acc0, w[0][2 * i], d2.x); ffma2(acc1, w[1][2 * i], d2.x);
        ffma2(acc2, w[2][2 * i], d2.x); ffma2(acc3, w[3][2 * i], d2.x);
        ffma2(acc0, w[0][2 * i + 1], d2.y); ffma2(acc1, w[1][2 * i + 1], d2.y);
        ffma2(acc2, w[2][2 * i + 1], d2.y); ffma2(acc3, w[3][2 * i + 1], d2.y);
      } else {
        float dummy; unpack2(d2.x, dtail, dummy);
      }
    }
    float l0, l1;
    unpack2(acc0, l0, l1); const float p0 = fmaf(wt[0], dtail, l0 + l1);
    unpack2(acc1, l0, l1); const float p1 = fmaf(wt[1], dtail, l0 + l1);
    unpack2(acc2, l0, l1); const float p2 = fmaf(wt[2], dtail, l0 + l1);
    unpack2(acc3, l0, l1); const float p3 = fmaf(wt[3], dtail, l0 + l1);
    return group_reduce4<RB>(p0, p1, p2, p3, hi, lo);
  };
  for (int u = 0; u < TA; u++) {
    {
      const int un = u + kStage - 1;
      if (pw && un < TA) stage(un, ncolA + (kStage - 1) * dt, TA, 0);
      if (pw && un < TB) stage(un, ncolB + (kStage - 1) * dt, TB, SLINE);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const unsigned sslot = (u & (kStage - 1)) * STG;
    const bool liveB = u < TB;
    const unsigned wsel = (unsigned)(u & 1) << 3, rsel = (unsigned)((u - 1) & 1) << 3;
    const unsigned par = ((unsigned)(u - 1) >> 1) & 1u;
    const bool sender = send && tid0 < Cfg::THREADS;
    if (tid0 == 0) {
      if (u + 1 < TA) mbar_arm_(bar0 + wsel, NO * 4);
      if (u + 1 < TB) mbar_arm_(bar0 + 16 + wsel, NO * 4);
    }
    // ---- line A
    float dlA = 0.f;
    if (pw) {
      if (u > 0) mbar_wait_(bar0 + rsel, par);
      dlA = pointwise(st_addr0 + sslot, pread, wr_addr, u + 1 == TA, dccA);
    }
    __syncthreads();
    const float partA = matvec(rd_addr);
    if (u + 1 < TA && sender) st_async_f32(pdst, bar0 + wsel, dst_rank, partA);
    if (pw) DGb[ncolA * ROWS + 4 * punit + pg] = dlA;
    ncolA += dt;
    // ---- line B
    float dlB = 0.f, partB = 0.f;
    if (liveB && pw) {
      if (u > 0) mbar_wait_(bar0 + 16 + rsel, par);
      dlB = pointwise(st_addr0 + SLINE + sslot, pread + PLINE, wr_addr + DLINE, u + 1 == TB, dccB);
    }
    __syncthreads();
    if (liveB) partB = matvec(rd_addr + DLINE);
    if (liveB && u + 1 < TB && sender) st_async_f32(pdst + PLINE, bar0 + 16 + wsel, dst_rank, partB);
    if (liveB) {
      if (pw) DGb[ncolB * ROWS + 4 * punit + pg] = dlB;
      ncolB += dt;
    }
    rd_addr += dtog; wr_addr += dtog; dtog = -dtog;
    pread = psrc + ((ptog > 0) ? 0u : PBUF);
    pdst += ptog; ptog = -ptog;
  }
  cluster_sync_();        // nobody leaves while a peer might still be using the cluster's shared memory windows
}

template <int NO> constexpr size_t bwd_cluster2_smem() {
  typedef ClCfg<NO> Cfg;
  return (size_t)(4 * (4 * Cfg::UC / Cfg::SL) * Cfg::SSTR + 4 * Cfg::CS * Cfg::UC + 2 * kStage * Cfg::TPAD * 8) * sizeof(float);
}

template <int NO> constexpr size_t bwd_cluster_smem() {
  typedef ClCfg<NO> Cfg;
  return (size_t)(2 * (4 * Cfg::UC / Cfg::SL) * Cfg::SSTR + 2 * Cfg::CS * Cfg::UC + kStage * Cfg::TPAD * 8) * sizeof(float);
}

// two lines per cluster (lstm_*_cluster2 with the barrier exchange, lstm_*_cluster4 with the mbarrier exchange):
// 1 always, 0 never, -1 (default): when there are more (line, direction) chains than resident clusters (CLSTM_B200_CLUSTER_PAIR)
int g_pair_mode = -1;
int g_fwd_mbar = 1;      // exchange through mbarrier + st.async (lstm_*_cluster3); 0: cluster barrier (CLSTM_B200_CLUSTER_MBAR)

template <int NO>
cudaError_t launch_cluster(bool fwd, cudaStream_t st, const Lines& ln, const void* args) {
  typedef ClCfg<NO> Cfg;
  cudaLaunchConfig_t cfg{};
  const int ndir = fwd ? static_cast<const LstmFwdArgs*>(args)->ndir : static_cast<const LstmBwdArgs*>(args)->ndir;
  // measured (B200, ms per pass fwd / bwd): cfg3 (4-CTA clusters) barrier 9.5 / 9.9, mbarrier 7.0 / 9.5;
  // cfg4 (16-CTA clusters) barrier 99.6 / 100.8, paired barrier 88.8 / 85.7, mbarrier 68.2 / 87.2 => mbarrier kernels by default
  // with the mbarrier exchange pairing wins for both cluster sizes: cfg3 20.9 -> 19.6 ms, cfg4 170.5 -> 159.2 ms per step
  const bool pair = g_pair_mode == 1 || (g_pair_mode < 0 && g_fwd_mbar && ln.B * ndir > 2 * (148 / Cfg::CS));
  cfg.gridDim = dim3((pair ? (ln.B + 1) / 2 : ln.B) * Cfg::CS, ndir, 1);
  cfg.blockDim = dim3(Cfg::TPAD, 1, 1);
  cfg.dynamicSmemBytes = fwd ? 0 : (pair ? bwd_cluster2_smem<NO>() : bwd_cluster_smem<NO>());
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg::CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!pair && g_fwd_mbar) {
    if (fwd) return cudaLaunchKernelEx(&cfg, lstm_fwd_cluster3<NO>, ln, *static_cast<const LstmFwdArgs*>(args));
    return cudaLaunchKernelEx(&cfg, lstm_bwd_cluster3<NO>, ln, *static_cast<const LstmBwdArgs*>(args));
  }
  if (pair && g_fwd_mbar) {
    if (fwd) return cudaLaunchKernelEx(&cfg, lstm_fwd_cluster4<NO>, ln, *static_cast<const LstmFwdArgs*>(args));
    return cudaLaunchKernelEx(&cfg, lstm_bwd_cluster4<NO>, ln, *static_cast<const LstmBwdArgs*>(args));
  }
  if (pair) {
    if (fwd) return cudaLaunchKernelEx(&cfg, lstm_fwd_cluster2<NO>, ln, *static_cast<const LstmFwdArgs*>(args));
    return cudaLaunchKernelEx(&cfg, lstm_bwd_cluster2<NO>, ln, *static_cast<const LstmBwdArgs*>(args));
  }
  if (fwd) return cudaLaunchKernelEx(&cfg, lstm_fwd_cluster<NO>, ln, *static_cast<const LstmFwdArgs*>(args));
  return cudaLaunchKernelEx(&cfg, lstm_bwd_cluster<NO>, ln, *static_cast<const LstmBwdArgs*>(args));
}

template <int NO>
cudaError_t configure_one() {
  cudaError_t e = cudaFuncSetAttribute(lstm_bwd_cluster<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bwd_cluster_smem<NO>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(lstm_bwd_cluster2<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_cluster2_smem<NO>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(lstm_bwd_cluster3<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_cluster_smem<NO>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(lstm_bwd_cluster4<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_cluster2_smem<NO>());
  if (e != cudaSuccess) return e;
  if (ClCfg<NO>::CS > 8) {
    e = cudaFuncSetAttribute(lstm_fwd_cluster<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_bwd_cluster<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_cluster2<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_cluster3<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_bwd_cluster3<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_cluster4<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_bwd_cluster4<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_bwd_cluster2<NO>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  }
  return e;
}

}  // namespace

bool lstm_cluster_supported(int no) { return no == 200 || no == 400; }

int lstm_cluster_configure() {
  if (const char* m = getenv("CLSTM_B200_CLUSTER_PAIR")) g_pair_mode = (m[0] == '1') ? 1 : (m[0] == '0' ? 0 : -1);
  else g_pair_mode = -1;
  if (const char* m = getenv("CLSTM_B200_CLUSTER_MBAR")) g_fwd_mbar = (m[0] != '0');
  cudaError_t e = configure_one<200>();
  if (e == cudaSuccess) e = configure_one<400>();
  return (int)e;
}

int lstm_cluster_forward(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  if (a.no == 200) return (int)launch_cluster<200>(true, st, ln, &a);
  if (a.no == 400) return (int)launch_cluster<400>(true, st, ln, &a);
  return -1;
}
int lstm_cluster_backward(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  if (a.no == 200) return (int)launch_cluster<200>(false, st, ln, &a);
  if (a.no == 400) return (int)launch_cluster<400>(false, st, ln, &a);
  return -1;
}

}  // namespace cb200
