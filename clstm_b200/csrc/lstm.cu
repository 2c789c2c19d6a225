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
  const int b = ln.order[blockIdx.x], d = blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
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
      const float act = ((r & 3) == 3) ? tanhf_(acc) : sigmoidf_(acc);
      act_s[r] = act;
      G[n * no4 + r] = act;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < no; j += blockDim.x) {
      const float gi = act_s[4 * j], gf = act_s[4 * j + 1], go = act_s[4 * j + 2], ci = act_s[4 * j + 3];
      float c = ci * gi;                       // forward_statemem clstm_compute.cc:504-508
      if (s > 0) c = fmaf(gf, c_s[j], c);
      c_s[j] = c;
      const float hh = tanhf_(c) * go;         // forward_nonlingate :530-537
      h_s[j] = hh;
      C[n * no + j] = c;
      a.H[n * (2 * no) + d * no + j] = hh;
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
  const int b = ln.order[blockIdx.x], d = blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
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
      const float dh = a.dH[n * (2 * no) + d * no + j] +
                       (part_s[j] + part_s[no + j] + part_s[2 * no + j] + part_s[3 * no + j]);
      const float4 g4 = *reinterpret_cast<const float4*>(G + n * no4 + 4 * j);
      const float gi = g4.x, gf = g4.y, go = g4.z, ci = g4.w;
      const float c = C[n * no + j];
      const float th = tanhf_(c);                        // backward_nonlingate clstm_compute.cc:539-547
      const float dgo = th * dh;
      const float dc = dcc_s[j] + (1.f - th * th) * (go * dh);
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
      o.w = (1.f - ci * ci) * dci;
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
template <int NO> struct RegCfg {
  static constexpr int NOP = (NO + 3) & ~3;            // K padded to a multiple of 4 (float4 smem reads)
  static constexpr int NP = NOP / 2;                   // packed pairs per thread
  static constexpr int ROWS = 4 * NO;
  static constexpr int THREADS = (ROWS + 31) & ~31;
};

// Forward.  thread r (< 4*NO) owns row r of R: pre[r] = XP[n][r] + sum_k R[r][k] h[k].
template <int NO>
__global__ void __launch_bounds__(RegCfg<NO>::THREADS, 1) lstm_fwd_regs(Lines ln, LstmFwdArgs a) {
  typedef RegCfg<NO> Cfg;
  constexpr int NOP = Cfg::NOP, NP = Cfg::NP, ROWS = Cfg::ROWS;
  __shared__ __align__(16) float h_s[2][NOP];
  const int tid = threadIdx.x;
  const int b = ln.order[blockIdx.x], d = blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const bool active = tid < ROWS;
  const int r = active ? tid : ROWS - 1;   // padding threads shadow the last row (results discarded)
  const int j = r >> 2, g = r & 3;
  const float* __restrict__ XP = a.XP[d] + r;
  float* __restrict__ G = a.G[d] + r;
  float* __restrict__ C = a.C[d] + j;
  float* __restrict__ Hp = a.Hprev[d] + j;
  float* __restrict__ H = a.H + d * NO + j;

  u64 w[NP];
  {
    const float* Rr = a.R[d] + (size_t)r * NO;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const float w0 = (2 * p < NO) ? Rr[2 * p] : 0.f;
      const float w1 = (2 * p + 1 < NO) ? Rr[2 * p + 1] : 0.f;
      w[p] = pack2(w0, w1);
    }
  }
  for (int k = tid; k < 2 * NOP; k += blockDim.x) (&h_s[0][0])[k] = 0.f;
  __syncthreads();

  const int dt = d ? -1 : 1;
  int t = d ? T - 1 : 0;
  // software prefetch of the input projection, 3 steps ahead
  float xp0 = 0.f, xp1 = 0.f, xp2 = 0.f;
  if (T > 0) xp0 = XP[(size_t)(off + t) * ROWS];
  if (T > 1) xp1 = XP[(size_t)(off + t + dt) * ROWS];
  if (T > 2) xp2 = XP[(size_t)(off + t + 2 * dt) * ROWS];
  float c = 0.f;
  for (int s = 0; s < T; s++, t += dt) {
    const size_t n = (size_t)off + t;
    float xp3 = 0.f;
    if (s + 3 < T) xp3 = XP[(size_t)(off + t + 3 * dt) * ROWS];
    const ulonglong2* hv = reinterpret_cast<const ulonglong2*>(h_s[s & 1]);
    u64 acc0 = pack2(xp0, 0.f), acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
#pragma unroll
    for (int q = 0; q < NOP / 4; q++) {
      const ulonglong2 h2 = hv[q];
      if (q & 1) { ffma2(acc2, w[2 * q], h2.x); ffma2(acc3, w[2 * q + 1], h2.y); }
      else       { ffma2(acc0, w[2 * q], h2.x); ffma2(acc1, w[2 * q + 1], h2.y); }
    }
    float s0, s1, s2, s3, s4, s5, s6, s7;
    unpack2(acc0, s0, s1); unpack2(acc1, s2, s3); unpack2(acc2, s4, s5); unpack2(acc3, s6, s7);
    const float pre = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
    // gates: sigmoid for gi,gf,go ; tanh for ci  (forward_full1, clstm.cc:614-617)
    const float sc = (g == 3) ? 2.f : 1.f;
    const float sg = 1.0f / (1.0f + expf(-sc * pre));
    const float act = (g == 3) ? 2.f * sg - 1.f : sg;
    if (active) G[n * ROWS] = act;
    const int base = (tid & 31) & ~3;
    const float gi = __shfl_sync(0xffffffffu, act, base + 0);
    const float gf = __shfl_sync(0xffffffffu, act, base + 1);
    const float go = __shfl_sync(0xffffffffu, act, base + 2);
    const float ci = __shfl_sync(0xffffffffu, act, base + 3);
    float cn = ci * gi;                                  // forward_statemem clstm_compute.cc:504-508
    if (s > 0) cn = fmaf(gf, c, cn);
    c = cn;
    const float hh = tanhf_(c) * go;                     // forward_nonlingate :530-537
    if (active) {
      if (g == 0) { h_s[(s + 1) & 1][j] = hh; H[n * (2 * NO)] = hh; }
      else if (g == 1) { C[n * NO] = c; }
      else if (g == 2) {
        if (s + 1 < T) Hp[(size_t)(off + t + dt) * NO] = hh;
      } else {
        if (s == 0) Hp[n * NO] = 0.f;
      }
    }
    xp0 = xp1; xp1 = xp2; xp2 = xp3;
    __syncthreads();
  }
}

// Backward.  thread (k = tid>>2, p = tid&3) owns R[4i+p][k], i < NO: the part of column k of R that belongs to
// gate p.  Per step: quad k recomputes the pointwise deltas of hidden unit k (redundantly in its 4 lanes), lane p
// publishes delta_p, then every thread accumulates its quarter of  dh_prev[k] = sum_r R[r][k] delta[r]  and the
// quad reduces with two shuffles.
template <int NO>
__global__ void __launch_bounds__(RegCfg<NO>::THREADS, 1) lstm_bwd_regs(Lines ln, LstmBwdArgs a) {
  typedef RegCfg<NO> Cfg;
  constexpr int NOP = Cfg::NOP, NP = Cfg::NP, ROWS = Cfg::ROWS;
  __shared__ __align__(16) float dg_s[2][4][NOP];        // [buffer][gate][unit]
  const int tid = threadIdx.x;
  const int b = ln.order[blockIdx.x], d = blockIdx.y;
  const int T = ln.T[b], off = ln.off[b];
  const bool active = tid < ROWS;
  const int rr = active ? tid : ROWS - 1;
  const int k = rr >> 2, p = rr & 3;
  const float* __restrict__ G = a.G[d] + 4 * k;
  const float* __restrict__ C = a.C[d] + k;
  const float* __restrict__ dH = a.dH + d * NO + k;
  float* __restrict__ DG = a.DG[d] + rr;

  u64 w[NP];
  {
    const float* R = a.R[d];
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const int i0 = 2 * q, i1 = 2 * q + 1;
      const float w0 = (i0 < NO) ? R[(size_t)(4 * i0 + p) * NO + k] : 0.f;
      const float w1 = (i1 < NO) ? R[(size_t)(4 * i1 + p) * NO + k] : 0.f;
      w[q] = pack2(w0, w1);
    }
  }
  for (int i = tid; i < 2 * 4 * NOP; i += blockDim.x) (&dg_s[0][0][0])[i] = 0.f;
  __syncthreads();

  // walk the direction's forward steps backwards: fs = T-1 .. 0 ; column t = d ? T-1-fs : fs
  const int dt = d ? 1 : -1;                 // column increment per backward step
  int t = d ? 0 : T - 1;
  float dhrec = 0.f, dcc = 0.f;
  // prefetch queue (1 step ahead): gates, cell, previous cell, upstream delta
  float4 g_n = make_float4(0.f, 0.f, 0.f, 0.f);
  float c_n = 0.f, dh_n = 0.f;
  if (T > 0) {
    const size_t n = (size_t)off + t;
    g_n = *reinterpret_cast<const float4*>(G + n * ROWS);
    c_n = C[n * NO];
    dh_n = dH[n * (2 * NO)];
  }
  int buf = 0;
  for (int fs = T - 1; fs >= 0; fs--, t += dt, buf ^= 1) {
    const size_t n = (size_t)off + t;
    const float4 g4 = g_n;
    const float c = c_n;
    const float dhu = dh_n;
    float cprev = 0.f;
    if (fs > 0) {                            // loads for the next backward step (forward step fs-1)
      const size_t n1 = (size_t)off + t + dt;
      g_n = *reinterpret_cast<const float4*>(G + n1 * ROWS);
      c_n = C[n1 * NO];
      dh_n = dH[n1 * (2 * NO)];
      cprev = c_n;   // consumed below only after the loads land; it is the same value as next step's c
    }
    const float gi = g4.x, gf = g4.y, go = g4.z, ci = g4.w;
    const float dh = dhu + dhrec;
    const float th = tanhf_(c);                              // backward_nonlingate clstm_compute.cc:539-547
    const float dgo = th * dh;
    const float dc = dcc + (1.f - th * th) * (go * dh);
    float dgf = 0.f;
    dcc = 0.f;
    if (fs > 0) {                                            // backward_statemem :509-515
      dgf = dc * cprev;
      dcc = dc * gf;
    }
    const float dgi = dc * ci, dci = dc * gi;
    float dl;                                                // backward_nonlin0 :231-267, this lane's gate p
    if (p == 0) dl = gi * (1.f - gi) * dgi;
    else if (p == 1) dl = gf * (1.f - gf) * dgf;
    else if (p == 2) dl = go * (1.f - go) * dgo;
    else dl = (1.f - ci * ci) * dci;
    if (active) {
      dg_s[buf][p][k] = dl;
      DG[n * ROWS] = dl;
    }
    __syncthreads();
    // quarter of dh_{fs-1}[k] = sum_i R[4i+p][k] * delta_p[i]
    const ulonglong2* dv = reinterpret_cast<const ulonglong2*>(dg_s[buf][p]);
    u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
#pragma unroll
    for (int q = 0; q < NOP / 4; q++) {
      const ulonglong2 d2 = dv[q];
      if (q & 1) { ffma2(acc2, w[2 * q], d2.x); ffma2(acc3, w[2 * q + 1], d2.y); }
      else       { ffma2(acc0, w[2 * q], d2.x); ffma2(acc1, w[2 * q + 1], d2.y); }
    }
    float s0, s1, s2, s3, s4, s5, s6, s7;
    unpack2(acc0, s0, s1); unpack2(acc1, s2, s3); unpack2(acc2, s4, s5); unpack2(acc3, s6, s7);
    float part = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    dhrec = part;
  }
}

template <int NO>
void launch_fwd_regs(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  lstm_fwd_regs<NO><<<dim3(ln.B, 2), RegCfg<NO>::THREADS, 0, st>>>(ln, a);
}
template <int NO>
void launch_bwd_regs(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  lstm_bwd_regs<NO><<<dim3(ln.B, 2), RegCfg<NO>::THREADS, 0, st>>>(ln, a);
}

bool has_regs_variant(int no) { return no == 16 || no == 32 || no == 50 || no == 64 || no == 100; }
int generic_threads(int no) {
  int th = 4 * no;
  th = (th + 31) & ~31;
  return th > 1024 ? 1024 : (th < 64 ? 64 : th);
}
}  // namespace

const char* lstm_variant_for(int no) { return has_regs_variant(no) ? "regs" : "generic"; }

int lstm_configure() {
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(lstm_bwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  return (int)e;
}

const char* lstm_forward(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  switch (a.no) {
    case 16: launch_fwd_regs<16>(st, ln, a); return "regs";
    case 32: launch_fwd_regs<32>(st, ln, a); return "regs";
    case 50: launch_fwd_regs<50>(st, ln, a); return "regs";
    case 64: launch_fwd_regs<64>(st, ln, a); return "regs";
    case 100: launch_fwd_regs<100>(st, ln, a); return "regs";
    default: break;
  }
  const size_t smem = (size_t)6 * a.no * sizeof(float);
  lstm_fwd_generic<<<dim3(ln.B, 2), generic_threads(a.no), smem, st>>>(ln, a);
  return "generic";
}

const char* lstm_backward(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  switch (a.no) {
    case 16: launch_bwd_regs<16>(st, ln, a); return "regs";
    case 32: launch_bwd_regs<32>(st, ln, a); return "regs";
    case 50: launch_bwd_regs<50>(st, ln, a); return "regs";
    case 64: launch_bwd_regs<64>(st, ln, a); return "regs";
    case 100: launch_bwd_regs<100>(st, ln, a); return "regs";
    default: break;
  }
  const size_t smem = (size_t)9 * a.no * sizeof(float);
  lstm_bwd_generic<<<dim3(ln.B, 2), generic_threads(a.no), smem, st>>>(ln, a);
  return "generic";
}

}  // namespace cb200
