// ctc.cu -- OCRopus-style CTC alignment, one CTA per text line.
//
// Restates /root/reference/ctc.cc:24-134 for one-hot targets built by mktargets
// (ctc.cc:148-157): S = 2L+1 states, even states = class 0 (blank), odd state s = transcript[(s-1)/2].
//   phase A  o' = max(1e-5,o) / sum ;  lmatch(t,s) = log o'(t, label_s)                       ctc.cc:68-77
//   phase B  forward lattice lr and the backward lattice rl (forward algorithm on the doubly flipped lmatch)
//            with the soft start penalty skip = -5 per state/step and log_add's |x-y|>10 cutoff ctc.cc:24-55
//   phase C  epath = limexp(lr + rl - max) ; normalised per STATE over TIME                    ctc.cc:83-89
//   phase D  aligned(t,c) = sum_{s: label_s = c} epath(t,s) ; normalised per t ; delta = aligned - out
//                                                                                   ctc.cc:92-109, clstmhl.h:211-212
// The two lattices are the only serial part: each is walked by ONE warp (lane l owns KS consecutive states), so a
// time step costs one warp shuffle and no block barrier; both warps run concurrently.  Everything else is
// parallel over (t, s).  Sums the reference accumulates in double are accumulated in double here too.
#include "kernels.h"

namespace cb200 {
namespace {

__device__ __forceinline__ float limexp_(float x) {  // tensor.h:78-82
  if (x < -30.f) return expf(-30.f);
  if (x > 30.f) return expf(30.f);
  return expf(x);
}
__device__ __forceinline__ float log_add_(float x, float y) {  // tensor.h:86-89
  if (fabsf(x - y) > 10.f) return fmaxf(x, y);
  return logf(expf(x - y) + 1.f) + y;
}

constexpr int CTC_THREADS = 256;

// One lattice pass by one warp.  rev=0: lr(t,s).  rev=1: processes i-th step on column T-1-i and state index jj
// on real state S-1-jj, result stored at rl(T-1-i, S-1-jj)  (forwardbackward, ctc.cc:42-55).
template <int KS>
__device__ void lattice_pass(const float* __restrict__ lm, float* __restrict__ out, int T, int S, bool rev) {
  const int lane = threadIdx.x & 31;
  float v[KS];
  float m_cur[KS], m_nxt[KS];
#pragma unroll
  for (int k = 0; k < KS; k++) {
    const int jj = lane * KS + k;
    v[k] = (float)(-5.0 * jj);                       // ctc.cc:30
    m_cur[k] = 0.f;
    m_nxt[k] = 0.f;
  }
  auto load_row = [&](int i, float* dst) {
    const int t = rev ? T - 1 - i : i;
    const float* row = lm + (size_t)t * S;
#pragma unroll
    for (int k = 0; k < KS; k++) {
      const int jj = lane * KS + k;
      if (jj < S) dst[k] = row[rev ? S - 1 - jj : jj];
    }
  };
  if (T > 0) load_row(0, m_cur);
  for (int i = 0; i < T; i++) {
    if (i + 1 < T) load_row(i + 1, m_nxt);
    float below = __shfl_up_sync(0xffffffffu, v[KS - 1], 1);   // old v(jj-1) of the first state of this lane
    if (lane == 0) below = (float)(-5.0 * i);                  // w(0) = skip*i   ctc.cc:32
#pragma unroll
    for (int k = KS - 1; k >= 0; k--) {
      const float w = (k == 0) ? below : v[k - 1];             // w(j) = v(j-1) before the update  ctc.cc:33
      const float same = v[k] + m_cur[k];
      const float next = w + m_cur[k];
      v[k] = log_add_(same, next);
    }
    const int t = rev ? T - 1 - i : i;
    float* orow = out + (size_t)t * S;
#pragma unroll
    for (int k = 0; k < KS; k++) {
      const int jj = lane * KS + k;
      if (jj < S) orow[rev ? S - 1 - jj : jj] = v[k];
    }
#pragma unroll
    for (int k = 0; k < KS; k++) m_cur[k] = m_nxt[k];
  }
}

__device__ void lattice_dispatch(const float* lm, float* out, int T, int S, bool rev) {
  const int ks = (S + 31) / 32;
  if (ks <= 1) lattice_pass<1>(lm, out, T, S, rev);
  else if (ks <= 2) lattice_pass<2>(lm, out, T, S, rev);
  else if (ks <= 4) lattice_pass<4>(lm, out, T, S, rev);
  else if (ks <= 8) lattice_pass<8>(lm, out, T, S, rev);
  else if (ks <= 16) lattice_pass<16>(lm, out, T, S, rev);
  else lattice_pass<32>(lm, out, T, S, rev);
}

__global__ void __launch_bounds__(CTC_THREADS) ctc_align_kernel(Lines ln, CtcArgs a) {
  __shared__ int lab_s[kCtcMaxStates];        // class of each state
  __shared__ float red_s[CTC_THREADS / 32];
  __shared__ float mx_s;
  __shared__ double acc_s[(CTC_THREADS / 32) * kCtcMaxClasses];
  const int b = ln.order[blockIdx.x];
  const int T = ln.T[b], off = ln.off[b], L = ln.L[b];
  const bool raw = a.raw != 0;                 // raw: ln.L holds S and labels hold one class per state (ctc.cc:136-146)
  const int S = raw ? L : 2 * L + 1;
  const int nc = a.nc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (S > kCtcMaxStates || nc > kCtcMaxClasses) {
    if (tid == 0) atomicExch(a.status, 1);
    return;
  }
  const int* lab = ln.labels + ln.lab_off[b];
  for (int s = tid; s < S; s += CTC_THREADS)
    lab_s[s] = raw ? lab[s] : ((s & 1) ? lab[(s - 1) >> 1] : 0);                            // mktargets ctc.cc:148-157
  __syncthreads();
  const float* __restrict__ out = a.out + (size_t)off * nc;
  float* __restrict__ lm = a.lmatch + ln.lat_off[b];
  float* __restrict__ lr = a.lr + ln.lat_off[b];
  float* __restrict__ rl = a.rl + ln.lat_off[b];

  // ---- phase A1: one thread per column: asum1 (tensor.h:337-342) is a sequential Float sum, keep its order.
  //      The sums are parked in the first T entries of rl (rl is not written before phase B).
  for (int t = tid; t < T; t += CTC_THREADS) {
    const float* o = out + (size_t)t * nc;
    float sum = 0.f;
    for (int i = 0; i < nc; i++) sum += fmaxf(1e-5f, o[i]);
    rl[t] = sum;
  }
  __syncthreads();
  // ---- phase A2: one warp per column: lmatch(t,s) = log(o'(t,label_s)); the blank log is computed once
  for (int t = warp; t < T; t += CTC_THREADS / 32) {
    const float* o = out + (size_t)t * nc;
    const float sum = rl[t];
    float lblank = 0.f;
    if (lane == 0) lblank = (float)log((double)(fmaxf(1e-5f, o[0]) / sum));   // double log, stored as Float ctc.cc:73-76
    lblank = __shfl_sync(0xffffffffu, lblank, 0);
    float* lrow = lm + (size_t)t * S;
    for (int s = lane; s < S; s += 32) {
      float v = lblank;
      if (raw || (s & 1)) v = (float)log((double)(fmaxf(1e-5f, o[lab_s[s]]) / sum));
      lrow[s] = v;
    }
  }
  __syncthreads();

  // ---- phase B: warp 0 = forward lattice, warp 1 = backward lattice
  if (warp == 0) lattice_dispatch(lm, lr, T, S, false);
  else if (warp == 1) lattice_dispatch(lm, rl, T, S, true);
  __syncthreads();

  // ---- phase C: both = lr + rl (kept in lr), global max, epath, per-state normalisation over time
  const int TS = T * S;
  float mx = -INFINITY;
  for (int i = tid; i < TS; i += CTC_THREADS) {
    const float v = lr[i] + rl[i];
    lr[i] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
  if (lane == 0) red_s[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red_s[0];
    for (int i = 1; i < CTC_THREADS / 32; i++) m = fmaxf(m, red_s[i]);
    mx_s = m;
  }
  __syncthreads();
  mx = mx_s;
  for (int i = tid; i < TS; i += CTC_THREADS) lr[i] = limexp_(lr[i] - mx);      // epath  ctc.cc:83
  __syncthreads();
  for (int s = warp; s < S; s += CTC_THREADS / 32) {                             // ctc.cc:84-89
    double tot = 0.0;
    for (int t = lane; t < T; t += 32) tot += (double)lr[(size_t)t * S + s];
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o2);
    tot = fmax(1e-9, tot);
    for (int t = lane; t < T; t += 32) {
      const size_t i = (size_t)t * S + s;
      lr[i] = (float)((double)lr[i] / tot);
    }
  }
  __syncthreads();

  // ---- phase D: project states onto classes (double accumulators like ctc.cc:96-103), normalise per column,
  //      emit delta.  One warp per column; per-warp class accumulators in shared memory.
  float* __restrict__ al = a.aligned + (size_t)off * nc;
  float* __restrict__ dl = a.delta + (size_t)off * nc;
  double* acc = acc_s + warp * kCtcMaxClasses;
  for (int t = warp; t < T; t += CTC_THREADS / 32) {
    const float* ep = lr + (size_t)t * S;
    for (int c = lane; c < nc; c += 32) acc[c] = 0.0;
    __syncwarp();
    if (raw) {
      for (int s = lane; s < S; s += 32) atomicAdd(&acc[lab_s[s]], (double)ep[s]);
      __syncwarp();
    } else {
      double blank = 0.0;                                  // all even states are class 0
      for (int s = 2 * lane; s < S; s += 64) blank += (double)ep[s];
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) blank += __shfl_xor_sync(0xffffffffu, blank, o2);
      for (int s = 2 * lane + 1; s < S; s += 64) atomicAdd(&acc[lab_s[s]], (double)ep[s]);
      __syncwarp();
      if (lane == 0) acc[0] += blank;
      __syncwarp();
    }
    double tot = 0.0;                                      // row total of the Float-rounded values  ctc.cc:104-109
    for (int c = lane; c < nc; c += 32) tot += (double)(float)acc[c];
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o2);
    tot = fmax(tot, 1e-9);
    const float* o = out + (size_t)t * nc;
    for (int c = lane; c < nc; c += 32) {
      const float v = (float)((double)(float)acc[c] / tot);
      al[(size_t)t * nc + c] = v;
      dl[(size_t)t * nc + c] = v - o[c];
    }
    __syncwarp();
  }
}

}  // namespace

void ctc_align(cudaStream_t st, const Lines& ln, const CtcArgs& a) {
  ctc_align_kernel<<<ln.B, CTC_THREADS, 0, st>>>(ln, a);
}

}  // namespace cb200
