// lstm.cu -- the serial part of the bidirectional NPLSTM: one persistent CTA per (text line, direction)
// walks the line's columns in the direction's own time order and keeps the recurrent weights and the running
// (h, c) state on chip for the whole sequence.
//
// Reference semantics restated here (paths relative to /root/reference):
//   GenericNPLSTM<SIG,TANH,TANH>::forward   clstm.cc:600-621   (stack_delay, 4x full1, statemem, nonlingate)
//   GenericNPLSTM::backward                 clstm.cc:622-653   (nonlingate, statemem, 4x full1, stack_delay)
//   Reversed / Parallel wiring              clstm.cc:458-479, 506-544  -> direction 1 simply walks t = T-1..0 and
//                                           writes its h into columns [no, 2no) of the shared H matrix, so the
//                                           sequence reversals and the concat cost no memory traffic at all.
// What is NOT done here (hoisted into dense products over all columns, gemm.cu): the input half of the gate
// pre-activations  W[:,1:1+ni] x_t + W[:,0]  (XP), the weight derivatives and the input deltas.
//
// Weight rows are gate-interleaved: row r = 4*j + g, g in {0:gi, 1:gf, 2:go, 3:ci}, so the four gates of hidden
// unit j live in four adjacent lanes of one warp and the cell update needs warp shuffles only.
//
// Two variants:
//   "regs"    : template on NO; each thread keeps one full row of R (forward) / one quarter column of R (backward)
//               in registers as packed f32x2 pairs and uses the Blackwell packed FFMA2 (fma.rn.f32x2);
//               one __syncthreads per timestep.
//   "generic" : any NO; R streamed from L2 every step.  Correctness fallback for sizes without an instantiation.
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
// acc += a * b on two packed fp32 lanes (Blackwell FFMA2)
__device__ __forceinline__ void ffma2(u64& acc, u64 a, u64 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// tanh via the same exponential: 2*sigmoid(2x) - 1 (abs. error ~1e-7, well inside the 1e-4 parity bar)
__device__ __forceinline__ float tanhf_(float x) { return 2.0f / (1.0f + expf(-2.0f * x)) - 1.0f; }

// cell variants (GenericNPLSTM<F,G,H>, clstm.cc:546-668): G = nonlinearity of the cell input ci, H = of the cell output
__device__ __forceinline__ bool cell_g_relu(int cell) { return cell >= 2; }
__device__ __forceinline__ int cell_h_kind(int cell) { return (cell == 0 || cell == 2) ? 2 : (cell == 4 ? 3 : 0); }   // 2 TANH, 3 RELU, 0 LIN
__device__ __forceinline__ float cell_h(int kind, float c) { return kind == 2 ? tanhf_(c) : (kind == 3 ? fmaxf(c, 0.f) : c); }
// derivative in terms of the OUTPUT y of the nonlinearity (backward_nonlin0, clstm_compute.cc:231-267)
__device__ __forceinline__ float cell_h_deriv(int kind, float y) { return kind == 2 ? 1.f - y * y : (kind == 3 ? (y > 0.f ? 1.f : 0.f) : 1.f); }

// --------------------------------------------------------------------------------------------------------
// generic kernels
// --------------------------------------------------------------------------------------------------------
// smem: h[no] | c[no] | act[4no]
__global__ void lstm_fwd_generic(Lines ln, LstmFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int no = a.no, no4 = 4 * a.no;
  float* h_s = sm;
  float* c_s = sm + no;
  float* act_s = sm + 2 * no;
  const int b = ln.order[blockIdx.x], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const bool g_relu = cell_g_relu(a.cell);
  const int hk = cell_h_kind(a.cell);
  const float* __restrict__ XP = a.XP[d];
  const float* __restrict__ Rt = a.Rt[d];
  float* __restrict__ G = a.G[d];
  float* __restrict__ C = a.C[d];
  float* __restrict__ Hp = a.Hprev[d];
  for (int j = threadIdx.x; j < no; j += blockDim.x) { h_s[j] = 0.f; c_s[j] = 0.f; }
  __syncthreads();
  for (int s = 0; s < T; s++) {
    const int t = d ? T - 1 - s : s;
    const size_t n = (size_t)off + t;
    for (int r = threadIdx.x; r < no4; r += blockDim.x) {
      float acc = XP[n * no4 + r];
      for (int k = 0; k < no; k++) acc = fmaf(Rt[(size_t)k * no4 + r], h_s[k], acc);
      const float act = ((r & 3) == 3) ? (g_relu ? fmaxf(acc, 0.f) : tanhf_(acc)) : sigmoidf_(acc);
      act_s[r] = act;
      G[n * no4 + r] = act;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < no; j += blockDim.x) {
      const float gi = act_s[4 * j], gf = act_s[4 * j + 1], go = act_s[4 * j + 2], ci = act_s[4 * j + 3];
      float c = ci * gi;                       // forward_statemem clstm_compute.cc:504-508
      if (s > 0) c = fmaf(gf, c_s[j], c);
      c_s[j] = c;
      const float hh = cell_h(hk, c) * go;     // forward_nonlingate :530-537
      h_s[j] = hh;
      C[n * no + j] = c;
      a.H[n * a.hstride + a.hoff[d] + j] = hh;
      if (s == 0) Hp[n * no + j] = 0.f;
      if (s + 1 < T) {
        const size_t n2 = (size_t)off + (d ? t - 1 : t + 1);
        Hp[n2 * no + j] = hh;
      }
    }
    __syncthreads();
  }
}

// smem: dg[4no] | part[4no] | dcc[no]
__global__ void lstm_bwd_generic(Lines ln, LstmBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int no = a.no, no4 = 4 * a.no;
  float* dg_s = sm;
  float* part_s = sm + no4;
  float* dcc_s = sm + 2 * no4;
  const int b = ln.order[blockIdx.x], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const bool g_relu = cell_g_relu(a.cell);
  const int hk = cell_h_kind(a.cell);
  const float* __restrict__ R = a.R[d];
  const float* __restrict__ G = a.G[d];
  const float* __restrict__ C = a.C[d];
  float* __restrict__ DG = a.DG[d];
  for (int j = threadIdx.x; j < no4; j += blockDim.x) part_s[j] = 0.f;
  for (int j = threadIdx.x; j < no; j += blockDim.x) dcc_s[j] = 0.f;
  __syncthreads();
  for (int fs = T - 1; fs >= 0; fs--) {      // fs = forward step index of this direction
    const int t = d ? T - 1 - fs : fs;
    const size_t n = (size_t)off + t;
    const size_t npv = (size_t)off + (d ? t + 1 : t - 1);  // column of forward step fs-1
    for (int j = threadIdx.x; j < no; j += blockDim.x) {
      const float dh = a.dH[n * a.hstride + a.hoff[d] + j] +
                       (part_s[j] + part_s[no + j] + part_s[2 * no + j] + part_s[3 * no + j]);
      const float4 g4 = *reinterpret_cast<const float4*>(G + n * no4 + 4 * j);
      const float gi = g4.x, gf = g4.y, go = g4.z, ci = g4.w;
      const float c = C[n * no + j];
      const float th = cell_h(hk, c);                    // backward_nonlingate clstm_compute.cc:539-547
      const float dgo = th * dh;
      const float dc = dcc_s[j] + cell_h_deriv(hk, th) * (go * dh);
      float dgf = 0.f, carry = 0.f;
      if (fs > 0) {                                      // backward_statemem :509-515
        dgf = dc * C[npv * no + j];
        carry = dc * gf;
      }
      dcc_s[j] = carry;
      const float dgi = dc * ci, dci = dc * gi;
      float4 o;                                          // backward_nonlin0 :231-267
      o.x = gi * (1.f - gi) * dgi;
      o.y = gf * (1.f - gf) * dgf;
      o.z = go * (1.f - go) * dgo;
      o.w = (g_relu ? (ci > 0.f ? 1.f : 0.f) : (1.f - ci * ci)) * dci;
      *reinterpret_cast<float4*>(dg_s + 4 * j) = o;
      *reinterpret_cast<float4*>(DG + n * no4 + 4 * j) = o;
    }
    __syncthreads();
    // source.d[ni:] = sum_g W_g[:,1+ni:]^T delta_g  (backward_lin1 :296) -> delta of h_{fs-1}
    for (int idx = threadIdx.x; idx < no4; idx += blockDim.x) {
      const int p = idx / no, k = idx - p * no;
      float acc = 0.f;
      const int r0 = p * no;
      for (int r = r0; r < r0 + no; r++) acc = fmaf(R[(size_t)r * no + k], dg_s[r], acc);
      part_s[idx] = acc;
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------------------------
// register-resident kernels
// --------------------------------------------------------------------------------------------------------
// Thread layout (forward): quad j = tid>>2 is hidden unit j, lane q = tid&3 of the quad owns, for all FOUR gate
// rows 4j..4j+3, the quarter k in [q*SL, (q+1)*SL) of the recurrent weights (4 x SL values, register resident,
// packed f32x2).  Per step a thread reads only ITS quarter of h (SLP/4 LDS.128 instead of NO/4), runs four
// independent FFMA2 chains (one per gate row), and the quad reduce-scatters the four partial sums with three
// shuffles so that lane q ends up with the full pre-activation of gate q.
template <int NO> struct RegCfg {
  static constexpr int SL = (NO + 3) / 4;              // k-slice length per lane
  static constexpr int SLP = (SL + 3) & ~3;            // padded to float4 granularity (pad reads hit zeros)
  static constexpr int NVEC = SLP / 4;                 // LDS.128 per step
  static constexpr int SSTR = SLP + 4 * (((SLP / 4) & 1) ? 0 : 1);  // slice stride: odd multiple of 16 B => conflict-free
  static constexpr int ROWS = 4 * NO;
  static constexpr int THREADS = (ROWS + 31) & ~31;
  // backward: k-groups of 4 outputs x 16 row-slices of SLR rows
  static constexpr int KG = (NO + 3) / 4;
  static constexpr int SLR = (ROWS + 15) / 16;
  static constexpr int SLRP = (SLR + 3) & ~3;
  static constexpr int NVEC_R = SLRP / 4;
  static constexpr int RSTR = SLRP + 4 * (((SLRP / 4) & 1) ? 0 : 1);
  static constexpr int THREADS_B = (16 * KG + 31) & ~31;
};


// fast gate math: sigmoid(x) = 1/(1 + 2^(-x*log2e)) with the MUFU approximations (rel. error ~1e-6, far inside the
// 1e-4 parity bar; tests/test_gpu_parity.py checks it against the libm-based oracle)
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
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// per-thread asynchronous global->shared staging (LDGSTS): keeps the streamed operands off the register
// scoreboards that the shared-memory loads of the recurrence wait on
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
constexpr int kStage = 4;   // depth of the per-thread staging rings

template <int NO>
__global__ void __launch_bounds__(RegCfg<NO>::THREADS, 1) lstm_fwd_regs(Lines ln, LstmFwdArgs a) {
  typedef RegCfg<NO> Cfg;
  constexpr int SL = Cfg::SL, NVEC = Cfg::NVEC, SSTR = Cfg::SSTR, ROWS = Cfg::ROWS, THREADS = Cfg::THREADS;
  constexpr int NPF = SL / 2, TAIL = SL & 1;
  __shared__ __align__(16) float h_s[2][4 * SSTR];
  __shared__ float xp_s[kStage][THREADS];                // per-thread ring of staged input projections
  const int tid = threadIdx.x;
  const int b = ln.order[blockIdx.x], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  // padding quads (THREADS > ROWS) are exact clones of the last quad: same loads, same stores, same values, so the
  // whole loop is branch-free straight-line code for every thread
  const int r = (tid < ROWS) ? tid : ROWS - 4 + (tid & 3);
  const int j = r >> 2, q = r & 3;
  const float* __restrict__ XPb = d ? a.XP[1] : a.XP[0];
  float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  float* __restrict__ Hpb = d ? a.Hprev[1] : a.Hprev[0];
  float* __restrict__ Hb = a.H + a.hoff[d];

  u64 w[4][NPF > 0 ? NPF : 1];                           // full pairs of the slice
  float wt[4];                                           // odd tail element of the slice (SL odd)
  {
    const float* Rd = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float* Rr = Rd + (size_t)(4 * j + g) * NO + q * SL;
#pragma unroll
      for (int p = 0; p < NPF; p++) {
        const float w0 = (q * SL + 2 * p < NO) ? Rr[2 * p] : 0.f;
        const float w1 = (q * SL + 2 * p + 1 < NO) ? Rr[2 * p + 1] : 0.f;
        w[g][p] = pack2(w0, w1);
      }
      wt[g] = (TAIL && q * SL + SL - 1 < NO) ? Rr[SL - 1] : 0.f;
    }
  }
  for (int k = tid; k < 2 * 4 * SSTR; k += blockDim.x) (&h_s[0][0])[k] = 0.f;

  const unsigned hs_base = (unsigned)__cvta_generic_to_shared(&h_s[0][0]);
  constexpr unsigned BUFB = 4 * SSTR * 4;                                  // bytes per h buffer
  unsigned rd_addr = hs_base + q * (SSTR * 4);                             // this lane's k-slice, buffer 0
  unsigned wr_addr = hs_base + BUFB + ((j / SL) * SSTR + (j % SL)) * 4;    // h[j] in buffer 1
  const unsigned xs_addr = (unsigned)__cvta_generic_to_shared(&xp_s[0][tid]);

  // per-lane output stream, one unconditional store per step: q=0 (and its duplicate q=3) -> H[n] = h_t,
  // q=1 -> C[n] = c_t, q=2 -> Hprev[n] = h_{t-1} (zero at the first step)
  const int dt = d ? -1 : 1;
  const int t0 = d ? T - 1 : 0;
  float* __restrict__ obase = (q == 0 || q == 3) ? Hb : (q == 1) ? Cb : Hpb;
  const unsigned ostride = (q == 0 || q == 3) ? (unsigned)a.hstride : NO;
  unsigned ncol = off + t0;

  // stage the input projection of the first kStage-1 steps
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T) cp_async4(xs_addr + u * (THREADS * 4), XPb + (size_t)(ncol + u * dt) * ROWS + r);
    cp_async_commit();
  }
  __syncthreads();

  float c = 0.f, hprev = 0.f;
  const int m_c = (q == 1) ? -1 : 0, m_p = (q == 2) ? -1 : 0, m_h = ~(m_c | m_p);
  const bool hi2 = (q & 2) != 0, hi1 = (q & 1) != 0;
  const float sc = (q == 3) ? -2.f * kLog2e : -kLog2e;
  int tog = (int)BUFB;
  for (int s = 0; s < T; s++) {
    {  // stage step s+kStage-1, then make sure step s has landed (own element only: no cross-thread dependency)
      const int sn = s + kStage - 1;
      if (sn < T) cp_async4(xs_addr + (sn & (kStage - 1)) * (THREADS * 4), XPb + (size_t)(ncol + (kStage - 1) * dt) * ROWS + r);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float htail = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; i++) {
      const ulonglong2 h2 = lds_v2u64(rd_addr + 16 * i);
      if (2 * i < NPF) {
        ffma2(acc0, w[0][2 * i], h2.x); ffma2(acc1, w[1][2 * i], h2.x);
        ffma2(acc2, w[2][2 * i], h2.x); ffma2(acc3, w[3][2 * i], h2.x);
      } else if (TAIL && 2 * i == NPF) {
        float dummy; unpack2(h2.x, htail, dummy);
      }
      if (2 * i + 1 < NPF) {
        ffma2(acc0, w[0][2 * i + 1], h2.y); ffma2(acc1, w[1][2 * i + 1], h2.y);
        ffma2(acc2, w[2][2 * i + 1], h2.y); ffma2(acc3, w[3][2 * i + 1], h2.y);
      } else if (TAIL && 2 * i + 1 == NPF) {
        float dummy; unpack2(h2.y, htail, dummy);
      }
    }
    const float xp = lds_f32(xs_addr + (s & (kStage - 1)) * (THREADS * 4));
    float lo, hi;
    unpack2(acc0, lo, hi); const float p0 = fmaf(wt[0], htail, lo + hi);
    unpack2(acc1, lo, hi); const float p1 = fmaf(wt[1], htail, lo + hi);
    unpack2(acc2, lo, hi); const float p2 = fmaf(wt[2], htail, lo + hi);
    unpack2(acc3, lo, hi); const float p3 = fmaf(wt[3], htail, lo + hi);
    // quad reduce-scatter: lane q ends with the total of gate row q
    float k0 = hi2 ? p2 : p0, k1 = hi2 ? p3 : p1;
    const float s0 = hi2 ? p0 : p2, s1 = hi2 ? p1 : p3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    float pre = hi1 ? k1 : k0;
    const float s2 = hi1 ? k0 : k1;
    pre += __shfl_xor_sync(0xffffffffu, s2, 1);
    pre += xp;
    // gates: sigmoid for gi,gf,go ; tanh(x) = 2*sigmoid(2x)-1 for ci  (forward_full1, clstm.cc:614-617)
    const float sg = rcp_approx(1.0f + ex2_approx(sc * pre));
    const float act = (q == 3) ? fmaf(2.f, sg, -1.f) : sg;
    Gb[ncol * ROWS + r] = act;
    const float gi = __shfl_sync(0xffffffffu, act, 0, 4);
    const float gf = __shfl_sync(0xffffffffu, act, 1, 4);
    const float go = __shfl_sync(0xffffffffu, act, 2, 4);
    const float ci = __shfl_sync(0xffffffffu, act, 3, 4);
    c = fmaf(gf, c, ci * gi);                            // forward_statemem clstm_compute.cc:504-508 (c = 0 before step 0)
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);
    const float hh = th * go;                            // forward_nonlingate :530-537
    sts_f32(wr_addr, hh);                                // all four lanes of the quad hold the same h: same address, same value
    obase[(size_t)ncol * ostride + j] = __int_as_float((__float_as_int(hh) & m_h) | (__float_as_int(c) & m_c) |
                                                        (__float_as_int(hprev) & m_p));   // branch-free per-lane select
    hprev = hh;
    ncol += dt;
    __syncthreads();
    rd_addr += tog; wr_addr -= tog; tog = -tog;          // swap the h double buffer
  }
}

// Backward.  The serial product is dh_prev[k] = sum_r R[r][k] delta[r] (r < 4*NO gate rows, k < NO).
// Thread tid: k-group kg = tid>>4 (outputs 4kg..4kg+3), row slice rs = tid&15 (rows rs*SLR .. +SLR): 4 x SLR weights
// in registers.  Per step it reads its slice of delta (NVEC_R LDS.128), runs four FFMA2 chains (one per output),
// and the 16 lanes of the k-group reduce with five shuffles, after which the four lanes {4kk..4kk+3} of the group
// all hold dh_prev[4kg+kk] -- i.e. thread tid holds dh_prev[tid>>2] and plays gate p = tid&3 in the pointwise part.
template <int NO>
__global__ void __launch_bounds__(RegCfg<NO>::THREADS_B, 1) lstm_bwd_regs(Lines ln, LstmBwdArgs a) {
  typedef RegCfg<NO> Cfg;
  constexpr int ROWS = Cfg::ROWS, SLR = Cfg::SLR, NVEC = Cfg::NVEC_R, RSTR = Cfg::RSTR, THREADS = Cfg::THREADS_B;
  constexpr int NPF = SLR / 2, TAIL = SLR & 1;
  constexpr bool EXACT = (NO % 4) == 0;                  // no padded outputs inside the last k-group
  extern __shared__ __align__(16) float bsm[];
  float* dg_s = bsm;                                     // [2][16 * RSTR]  delta double buffer, [row slice][pos]
  float* st_s = bsm + 2 * 16 * RSTR;                     // [kStage][THREADS][8] per-thread staging ring
  int tid = threadIdx.x;
  const int b = ln.order[blockIdx.x], d = a.d0 + blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  // padding 16-lane groups (THREADS > 16*KG) are exact clones of the last k-group => straight-line code
  const unsigned st_addr0 = (unsigned)__cvta_generic_to_shared(st_s + (size_t)tid * 8);
  if (tid >= 16 * Cfg::KG) tid -= 16;
  const int kraw = tid >> 2, p = tid & 3;
  const bool active = EXACT || kraw < NO;
  const int k = active ? kraw : NO - 1;
  const int kg4 = (tid >> 4) * 4, rs = tid & 15;
  const float* __restrict__ Gb = d ? a.G[1] : a.G[0];
  const float* __restrict__ Cb = d ? a.C[1] : a.C[0];
  const float* __restrict__ dHb = a.dH + a.hoff[d];
  float* __restrict__ DGb = d ? a.DG[1] : a.DG[0];
  const int row = 4 * k + p;                             // the delta row this thread publishes

  u64 w[4][NPF > 0 ? NPF : 1];
  float wt[4];
  {
    const float* R = d ? a.R[1] : a.R[0];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const int kc = kg4 + kk;
#pragma unroll
      for (int q = 0; q < NPF; q++) {
        const int r0 = rs * SLR + 2 * q, r1 = r0 + 1;
        const float w0 = (r0 < ROWS && kc < NO) ? R[(size_t)r0 * NO + kc] : 0.f;
        const float w1 = (r1 < ROWS && kc < NO) ? R[(size_t)r1 * NO + kc] : 0.f;
        w[kk][q] = pack2(w0, w1);
      }
      const int rt = rs * SLR + SLR - 1;
      wt[kk] = (TAIL && rt < ROWS && kc < NO) ? R[(size_t)rt * NO + kc] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < 2 * 16 * RSTR; i += blockDim.x) dg_s[i] = 0.f;

  const unsigned ds_base = (unsigned)__cvta_generic_to_shared(dg_s);
  constexpr unsigned BUFB = 16 * RSTR * 4;
  unsigned rd_addr = ds_base + rs * (RSTR * 4);                                  // this lane's row slice
  unsigned wr_addr = ds_base + ((row / SLR) * RSTR + (row % SLR)) * 4;           // where delta[row] lives
  int tog = (int)BUFB;
  constexpr unsigned STG = THREADS * 32;                                         // bytes per ring stage

  const int dt = d ? 1 : -1;                 // column increment per backward step
  unsigned ncol = off + (d ? 0 : T - 1);
  // staged record of a step: {gi,gf,go,ci | c, c_prev, dH, -}; c_prev = cell of the previous FORWARD step
  auto stage = [&](int u, unsigned col) {   // u = backward step index (0 .. T-1)
    const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
    cp_async16(sa, Gb + (size_t)col * ROWS + 4 * k);
    cp_async4(sa + 16, Cb + (size_t)col * NO + k);
    if (u + 1 < T) cp_async4(sa + 20, Cb + (size_t)(col + dt) * NO + k);
    cp_async4(sa + 24, dHb + (size_t)col * a.hstride + k);
  };
#pragma unroll
  for (int u = 0; u < kStage - 1; u++) {
    if (u < T) stage(u, ncol + u * dt);
    cp_async_commit();
  }
  __syncthreads();

  float dhrec = 0.f, dcc = 0.f;
  const bool b8 = (rs & 8) != 0, b4 = (rs & 4) != 0;
  const bool p_lo = (p & 1) != 0, p_hi = (p & 2) != 0;
  for (int u = 0; u < T; u++) {              // u-th backward step = forward step fs = T-1-u
    {
      const int un = u + kStage - 1;
      if (un < T) stage(un, ncol + (kStage - 1) * dt);
      cp_async_commit();
      cp_async_wait<kStage - 1>();
    }
    const bool first = (u + 1 == T);         // forward step 0: no previous cell, no forget-gate derivative
    const unsigned sa = st_addr0 + (u & (kStage - 1)) * STG;
    const ulonglong2 gq = lds_v2u64(sa);
    const ulonglong2 cq = lds_v2u64(sa + 16);
    float gi, gf, go, ci, c, cprev, dhu, unused;
    unpack2(gq.x, gi, gf); unpack2(gq.y, go, ci);
    unpack2(cq.x, c, cprev); unpack2(cq.y, dhu, unused);
    if (first) cprev = 0.f;
    const float dh = dhu + dhrec;
    const float th = fmaf(2.f, rcp_approx(1.0f + ex2_approx(-2.f * kLog2e * c)), -1.f);   // tanh(c) as in forward
    // backward_nonlingate clstm_compute.cc:539-547, backward_statemem :509-515, backward_nonlin0 :231-267
    const float dc = fmaf(1.f - th * th, go * dh, dcc);
    dcc = first ? 0.f : dc * gf;
    // this lane's gate p: value y, its incoming derivative e = A*B, and f'(y); two-level selects, no branches
    const float y0 = p_lo ? gf : gi, y1 = p_lo ? ci : go;
    const float y = p_hi ? y1 : y0;
    const float B0 = p_lo ? cprev : ci, B1 = p_lo ? gi : dh;
    const float Bv = p_hi ? B1 : B0;
    const float Av = (p == 2) ? th : dc;
    const float e = Av * Bv;
    const float fp = (1.f - y) * ((p == 3) ? (1.f + y) : y);
    const float dl = fp * e;
    if (active) {
      sts_f32(wr_addr, dl);
      DGb[ncol * ROWS + row] = dl;
    }
    __syncthreads();
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
    float dtail = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; i++) {
      const ulonglong2 d2 = lds_v2u64(rd_addr + 16 * i);
      if (2 * i < NPF) {
        ffma2(acc0, w[0][2 * i], d2.x); ffma2(acc1, w[1][2 * i], d2.x);
        ffma2(acc2, w[2][2 * i], d2.x); ffma2(acc3, w[3][2 * i], d2.x);
      } else if (TAIL && 2 * i == NPF) {
        float dummy; unpack2(d2.x, dtail, dummy);
      }
      if (2 * i + 1 < NPF) {
        ffma2(acc0, w[0][2 * i + 1], d2.y); ffma2(acc1, w[1][2 * i + 1], d2.y);
        ffma2(acc2, w[2][2 * i + 1], d2.y); ffma2(acc3, w[3][2 * i + 1], d2.y);
      } else if (TAIL && 2 * i + 1 == NPF) {
        float dummy; unpack2(d2.y, dtail, dummy);
      }
    }
    float lo, hi;
    unpack2(acc0, lo, hi); const float p0 = fmaf(wt[0], dtail, lo + hi);
    unpack2(acc1, lo, hi); const float p1 = fmaf(wt[1], dtail, lo + hi);
    unpack2(acc2, lo, hi); const float p2 = fmaf(wt[2], dtail, lo + hi);
    unpack2(acc3, lo, hi); const float p3 = fmaf(wt[3], dtail, lo + hi);
    // 16-lane reduce: scatter over bits 8 and 4 (lane group rs>>2 keeps output kk = rs>>2), then all-reduce bits 2,1
    float k0 = b8 ? p2 : p0, k1 = b8 ? p3 : p1;
    const float s0 = b8 ? p0 : p2, s1 = b8 ? p1 : p3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 8);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    float part = b4 ? k1 : k0;
    const float s2 = b4 ? k0 : k1;
    part += __shfl_xor_sync(0xffffffffu, s2, 4);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    dhrec = part;
    ncol += dt;
    rd_addr += tog; wr_addr += tog; tog = -tog;          // swap the delta double buffer
  }
}

template <int NO> constexpr size_t bwd_regs_smem() {
  return (size_t)(2 * 16 * RegCfg<NO>::RSTR + kStage * RegCfg<NO>::THREADS_B * 8) * sizeof(float);
}

template <int NO>
void launch_fwd_regs(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  lstm_fwd_regs<NO><<<dim3(ln.B, a.ndir), RegCfg<NO>::THREADS, 0, st>>>(ln, a);
}
template <int NO>
void launch_bwd_regs(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  lstm_bwd_regs<NO><<<dim3(ln.B, a.ndir), RegCfg<NO>::THREADS_B, bwd_regs_smem<NO>(), st>>>(ln, a);
}

bool has_regs_variant(int no) { return no == 16 || no == 32 || no == 50 || no == 64 || no == 100; }
int generic_threads(int no) {
  int th = 4 * no;
  th = (th + 31) & ~31;
  return th > 1024 ? 1024 : (th < 64 ? 64 : th);
}
}  // namespace

const char* lstm_variant_for(int no) {
  return has_regs_variant(no) ? "regs" : (lstm_cluster_supported(no) ? "cluster" : "generic");
}

int lstm_configure() {
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(lstm_bwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e != cudaSuccess) return (int)e;
#define CB200_BWD_SMEM(N_)                                                                                  \
  e = cudaFuncSetAttribute(lstm_bwd_regs<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_regs_smem<N_>()); \
  if (e != cudaSuccess) return (int)e;
  CB200_BWD_SMEM(16) CB200_BWD_SMEM(32) CB200_BWD_SMEM(50) CB200_BWD_SMEM(64) CB200_BWD_SMEM(100)
#undef CB200_BWD_SMEM
  if (e != cudaSuccess) return (int)e;
  return lstm_cluster_configure();
}

void lstm_forward_generic(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  lstm_fwd_generic<<<dim3(ln.B, a.ndir), generic_threads(a.no), (size_t)6 * a.no * sizeof(float), st>>>(ln, a);
}
void lstm_backward_generic(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  lstm_bwd_generic<<<dim3(ln.B, a.ndir), generic_threads(a.no), (size_t)9 * a.no * sizeof(float), st>>>(ln, a);
}

const char* lstm_forward(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  switch (a.cell == 0 ? a.no : -1) {        // the register / cluster kernels hard-wire the NPLSTM nonlinearities
    case 16: launch_fwd_regs<16>(st, ln, a); return "regs";
    case 32: launch_fwd_regs<32>(st, ln, a); return "regs";
    case 50: launch_fwd_regs<50>(st, ln, a); return "regs";
    case 64: launch_fwd_regs<64>(st, ln, a); return "regs";
    case 100: launch_fwd_regs<100>(st, ln, a); return "regs";
    default: break;
  }
  if (a.cell == 0 && lstm_cluster_supported(a.no)) {   // register-resident over a thread-block cluster (DSMEM exchange of h)
    if (lstm_cluster_forward(st, ln, a) == 0) return "cluster";
    cudaGetLastError();                 // cluster not schedulable on this device: stream the weights instead
  }
  const size_t smem = (size_t)6 * a.no * sizeof(float);
  lstm_fwd_generic<<<dim3(ln.B, a.ndir), generic_threads(a.no), smem, st>>>(ln, a);
  return "generic";
}

const char* lstm_backward(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  switch (a.cell == 0 ? a.no : -1) {
    case 16: launch_bwd_regs<16>(st, ln, a); return "regs";
    case 32: launch_bwd_regs<32>(st, ln, a); return "regs";
    case 50: launch_bwd_regs<50>(st, ln, a); return "regs";
    case 64: launch_bwd_regs<64>(st, ln, a); return "regs";
    case 100: launch_bwd_regs<100>(st, ln, a); return "regs";
    default: break;
  }
  if (a.cell == 0 && lstm_cluster_supported(a.no)) {
    if (lstm_cluster_backward(st, ln, a) == 0) return "cluster";
    cudaGetLastError();
  }
  const size_t smem = (size_t)9 * a.no * sizeof(float);
  lstm_bwd_generic<<<dim3(ln.B, a.ndir), generic_threads(a.no), smem, st>>>(ln, a);
  return "generic";
}

}  // namespace cb200
