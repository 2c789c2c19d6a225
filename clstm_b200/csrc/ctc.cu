// ctc.cu -- OCRopus-style CTC alignment on the device.
//
// Restates /root/reference/ctc.cc:24-134 for one-hot targets built by mktargets
// (ctc.cc:148-157: S = 2L+1 states, even states = class 0, odd state s = transcript[(s-1)/2]) or given explicitly as
// one class per state (the `Classes& targets` overload, ctc.cc:136-146; "raw" mode, used by the reference's own
// known-answer tests).  Four kernels; only the lattice recursion is serial:
//   ctc_lmatch     (grid: 32-column tiles)  o' = max(1e-5,o)/sum ; lmatch(t,s) = log o'(t, class_s)      ctc.cc:68-77
//   ctc_lattice    (grid: lines, 2 warps)   forward lattice lr and backward lattice rl, skip=-5 soft start,
//                                           log_add with the |x-y|>10 cutoff                           ctc.cc:24-55
//   ctc_max/ctc_sum (grid: lines x 8 slices) both = lr+rl ; max ; epath = limexp(both-max) ; per-STATE totals over TIME
//                                                                                                       ctc.cc:83-89
//   ctc_posterior  (grid: 32-column tiles)  epath /= total_s ; aligned(t,c) = sum_{s: class_s=c} epath(t,s) ;
//                                           aligned /= row total ; delta = aligned - out   ctc.cc:84-109, clstmhl.h:211
// Everything the reference accumulates in double is accumulated in double here; its Float sums keep their
// sequential order where that is free (asum1).  Lattices live in L2-resident global scratch (T x S per line).
#include "kernels.h"

namespace cb200 {
namespace {

__device__ __forceinline__ float limexp_(float x) {  // tensor.h:78-82
  if (x < -30.f) return expf(-30.f);
  if (x > 30.f) return expf(30.f);
  return expf(x);
}
// log_add for the lattice recursion: same branch structure as tensor.h:86-89, exp/log through the MUFU
// approximations (abs. error ~1e-7, below the Float resolution of lattice values, which run to O(10^2..10^3)).
__device__ __forceinline__ float log_add_fast(float x, float y) {
  const float d = x - y;
  float e, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d * 1.4426950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e + 1.f));
  const float r = fmaf(l, 0.6931471805599453f, y);
  return (fabsf(d) > 10.f) ? fmaxf(x, y) : r;
}
__device__ __forceinline__ int state_class(const int* lab, int s, bool raw) {   // mktargets ctc.cc:148-157
  return raw ? lab[s] : ((s & 1) ? lab[(s - 1) >> 1] : 0);
}

constexpr int TC = 32;            // columns per tile
constexpr int TILE_THREADS = 256;

// ------------------------------------------------------------------------------------------------ lmatch
// dynamic smem: tile[TC][ncp] | sum[TC] | lblank[TC] | cls[S]
__global__ void __launch_bounds__(TILE_THREADS) ctc_lmatch_kernel(Lines ln, CtcArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = ln.tile_line[blockIdx.x], t0 = ln.tile_t0[blockIdx.x];
  const int T = ln.T[b], off = ln.off[b], L = ln.L[b];
  const bool raw = a.raw != 0;
  const int S = raw ? L : 2 * L + 1;
  const int nc = a.nc, ncp = nc | 1;                     // odd row stride: conflict-free column walks
  const int ncols = min(TC, T - t0);
  float* tile = sm;
  float* sum_s = sm + TC * ncp;
  float* lbl_s = sum_s + TC;
  int* cls_s = reinterpret_cast<int*>(lbl_s + TC);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* lab = ln.labels + ln.lab_off[b];
  for (int s = tid; s < S; s += TILE_THREADS) cls_s[s] = state_class(lab, s, raw);
  const float* __restrict__ src = a.out + (size_t)(off + t0) * nc;
  for (int i = tid; i < ncols * nc; i += TILE_THREADS) {  // coalesced tile load
    const int c = i / nc, k = i - c * nc;
    tile[c * ncp + k] = fmaxf(1e-5f, src[i]);             // out(i) = fmax(lo, outputs)  ctc.cc:70
  }
  __syncthreads();
  if (warp == 0 && lane < ncols) {                        // asum1 (tensor.h:337-342): sequential Float sum
    const float* row = tile + lane * ncp;
    float sum = 0.f;
    for (int i = 0; i < nc; i++) sum += row[i];
    sum_s[lane] = sum;
    lbl_s[lane] = (float)log((double)(row[0] / sum));     // blank state: computed once per column
  }
  __syncthreads();
  float* __restrict__ lm = a.lmatch + ln.lat_off[b] + (size_t)t0 * S;
  for (int i = tid; i < ncols * S; i += TILE_THREADS) {   // memory order == idx order: coalesced stores
    const int c = i / S, s = i - c * S;
    float v;
    if (!raw && !(s & 1)) v = lbl_s[c];
    else v = (float)log((double)(tile[c * ncp + cls_s[s]] / sum_s[c]));   // double log stored as Float ctc.cc:73-76
    lm[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ lattice
// One lattice pass by W warps (one direction of one line).  rev=0: lr(t,s).  rev=1: processes i-th step on column T-1-i and
// state index jj on real state S-1-jj, result stored at rl(T-1-i, S-1-jj)  (forwardbackward, ctc.cc:42-55).
// Thread tg of the group owns KS consecutive states; a time step is one shuffle plus KS log_adds -- every state's update is the
// same expression whatever the distribution, so the lattices are bit-identical for every W.  W = 1 (short transcripts): a single
// warp has nothing to hide latency with, so the loop body is straight-line: running row pointers, a running -5*i, and lmatch
// rows prefetched PF steps ahead into a register ring without bounds checks (the lattice buffers are padded by kLatPad floats
// on both sides, so the prefetch may run past the line's rows).  W = 2 / 4 (long transcripts: a lane of a single warp would own
// up to 32 states = 64 MUFU operations per step on ONE scheduler): the warps of a direction sit on different schedulers; the
// value that crosses a warp boundary (old v of the previous warp's last state) goes through a double-buffered shared-memory
// slot and ONE named barrier per step.
template <int KS, int PF, int W>
__device__ void lattice_pass(const float* __restrict__ lm, float* __restrict__ out, int T, int S, bool rev, float* xch, int bar_id) {
  const int lane = threadIdx.x & 31;
  const int wg = (threadIdx.x >> 5) % W;                                   // warp inside the direction's group
  const int j0 = (wg * 32 + lane) * KS;
  const int dirk = rev ? -1 : 1;
  const long long rowstep = rev ? -(long long)S : (long long)S;
  const long long first = (rev ? (long long)(T - 1) * S + (S - 1 - j0) : (long long)j0);
  const float* lp = lm + first;
  float* op = out + first;
  bool ok[KS];
  float v[KS], mq[PF][KS];
#pragma unroll
  for (int k = 0; k < KS; k++) {
    ok[k] = (j0 + k) < S;
    v[k] = -5.f * (float)(j0 + k);                                        // ctc.cc:30 (exact in Float)
  }
#pragma unroll
  for (int u = 0; u < PF; u++)
#pragma unroll
    for (int k = 0; k < KS; k++) mq[u][k] = ok[k] ? lp[u * rowstep + k * dirk] : 0.f;
  const float* lpn = lp + PF * rowstep;
  float skipv = 0.f;                                                      // w(0) = skip*i   ctc.cc:32
  int par = 0;
  auto step = [&](float* m) {
    if (W > 1) {                                                          // old v of my last state -> the next warp's first state
      if (lane == 31) xch[par * W + wg] = v[KS - 1];
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * W) : "memory");
    }
    float below = __shfl_up_sync(0xffffffffu, v[KS - 1], 1);              // old v(jj-1) of this lane's first state
    if (lane == 0) below = (W == 1 || wg == 0) ? skipv : xch[par * W + wg - 1];
    par ^= 1;
    skipv -= 5.f;
#pragma unroll
    for (int k = KS - 1; k >= 0; k--) {
      const float w = (k == 0) ? below : v[k - 1];                        // w(j) = v(j-1) before the update  ctc.cc:33
      v[k] = log_add_fast(v[k] + m[k], w + m[k]);
    }
#pragma unroll
    for (int k = 0; k < KS; k++)
      if (ok[k]) op[k * dirk] = v[k];
    op += rowstep;
  };
  const int nfull = T / PF;
  for (int g = 0; g < nfull; g++) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      step(mq[u]);
#pragma unroll
      for (int k = 0; k < KS; k++) mq[u][k] = ok[k] ? lpn[k * dirk] : 0.f;   // row of step (g+1)*PF + u
      lpn += rowstep;
    }
  }
  const int rem = T - nfull * PF;
#pragma unroll
  for (int u = 0; u < PF; u++)
    if (u < rem) step(mq[u]);
}

constexpr int kLatWarps = 4;                     // warps per direction the launch provides
// warps per direction actually used for a transcript of S states (ks = states per lane of a single warp)
__device__ __forceinline__ int lattice_width(int S) {
  const int ks = (S + 31) / 32;
  return ks <= 2 ? 1 : (ks <= 4 ? 2 : 4);
}
template <int W>
__device__ void lattice_dispatch(const float* lm, float* out, int T, int S, bool rev, float* xch, int bar_id) {
  const int ks = (S + 32 * W - 1) / (32 * W);
  if (ks <= 1) lattice_pass<1, 8, W>(lm, out, T, S, rev, xch, bar_id);
  else if (ks <= 2) lattice_pass<2, 8, W>(lm, out, T, S, rev, xch, bar_id);
  else if (ks <= 4) lattice_pass<4, 4, W>(lm, out, T, S, rev, xch, bar_id);
  else lattice_pass<8, 2, W>(lm, out, T, S, rev, xch, bar_id);        // S <= kCtcMaxStates = 1024 = 4 warps x 32 lanes x 8
}

__global__ void __launch_bounds__(64 * kLatWarps) ctc_lattice_kernel(Lines ln, CtcArgs a) {
  __shared__ float xch_s[2][2 * kLatWarps];      // [direction][parity x warp]
  const int b = ln.order[blockIdx.x];
  const int T = ln.T[b], L = ln.L[b];
  const int S = a.raw ? L : 2 * L + 1;
  const float* lm = a.lmatch + ln.lat_off[b];
  const int W = lattice_width(S);
  const int warp = threadIdx.x >> 5;
  if (warp >= 2 * W) return;                     // (the named barriers below only count the 32 W threads of a direction)
  const bool rev = warp >= W;
  float* out = (rev ? a.rl : a.lr) + ln.lat_off[b];
  float* xch = xch_s[rev ? 1 : 0];
  const int bar_id = rev ? 2 : 1;
  if (W == 1) lattice_dispatch<1>(lm, out, T, S, rev, xch, bar_id);
  else if (W == 2) lattice_dispatch<2>(lm, out, T, S, rev, xch, bar_id);
  else lattice_dispatch<4>(lm, out, T, S, rev, xch, bar_id);
}

// ------------------------------------------------------------------------------------------------ stats
// Two kernels, grid (lines, kStatSlices): each CTA takes a slice of the line's time axis.
//   ctc_max:  slice maximum of both = lr + rl                                   (ctc.cc:54, amax2 :83)
//   ctc_sum:  lmatch <- epath = limexp(both - max) ; per-state partial sums over the slice's time range (double)
// The posterior kernel adds the kStatSlices partials in a fixed order: tot[s] = max(1e-9, sum_t epath(t,s)) (:84-87).
constexpr int kStatSlices = 8;
__global__ void __launch_bounds__(256) ctc_max_kernel(Lines ln, CtcArgs a) {
  __shared__ float red_s[8];
  const int b = ln.order[blockIdx.x], z = blockIdx.y;
  const int T = ln.T[b], L = ln.L[b];
  const int S = a.raw ? L : 2 * L + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ lr = a.lr + ln.lat_off[b];
  const float* __restrict__ rl = a.rl + ln.lat_off[b];
  const int rows = (T + kStatSlices - 1) / kStatSlices;
  const int i0 = min(T, z * rows) * S, i1 = min(T, (z + 1) * rows) * S;
  float mx = -INFINITY;
#pragma unroll 4
  for (int i = i0 + tid; i < i1; i += 256) mx = fmaxf(mx, lr[i] + rl[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red_s[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red_s[0];
    for (int i = 1; i < 8; i++) m = fmaxf(m, red_s[i]);
    a.mx_part[(size_t)b * kStatSlices + z] = m;
  }
}
__global__ void __launch_bounds__(256) ctc_sum_kernel(Lines ln, CtcArgs a) {
  __shared__ double part_s[4][64];
  const int b = ln.order[blockIdx.x], z = blockIdx.y;
  const int T = ln.T[b], L = ln.L[b];
  const int S = a.raw ? L : 2 * L + 1;
  const int tid = threadIdx.x;
  const float* __restrict__ lr = a.lr + ln.lat_off[b];
  const float* __restrict__ rl = a.rl + ln.lat_off[b];
  float* __restrict__ ep = a.lmatch + ln.lat_off[b];      // lmatch is dead after the lattice passes
  float mx = a.mx_part[(size_t)b * kStatSlices];
#pragma unroll
  for (int i = 1; i < kStatSlices; i++) mx = fmaxf(mx, a.mx_part[(size_t)b * kStatSlices + i]);
  const int rows = (T + kStatSlices - 1) / kStatSlices;
  const int t0 = min(T, z * rows), t1 = min(T, (z + 1) * rows);
  double* __restrict__ tot = a.tot + ((size_t)ln.st_off[b]) * kStatSlices + (size_t)z * S;
  const int sl = tid & 63, tr = tid >> 6;                 // 64 states x 4 time phases per sweep
  for (int s0 = 0; s0 < S; s0 += 64) {
    const int s = s0 + sl;
    double acc = 0.0;
    if (s < S) {
#pragma unroll 4
      for (int t = t0 + tr; t < t1; t += 4) {
        const int i = t * S + s;
        const float e = limexp_((lr[i] + rl[i]) - mx);    // epath  ctc.cc:83
        ep[i] = e;
        acc += (double)e;
      }
    }
    part_s[tr][sl] = acc;
    __syncthreads();
    if (tr == 0 && s < S) tot[s] = (part_s[0][sl] + part_s[1][sl]) + (part_s[2][sl] + part_s[3][sl]);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ posterior
// dynamic smem: acc[8][nc] doubles | tot[S] doubles | cls[S] ints
__global__ void __launch_bounds__(TILE_THREADS) ctc_posterior_kernel(Lines ln, CtcArgs a) {
  extern __shared__ __align__(16) double dsm[];
  const int b = ln.tile_line[blockIdx.x], t0 = ln.tile_t0[blockIdx.x];
  const int T = ln.T[b], off = ln.off[b], L = ln.L[b];
  const bool raw = a.raw != 0;
  const int S = raw ? L : 2 * L + 1;
  const int nc = a.nc;
  const int ncols = min(TC, T - t0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* acc = dsm + warp * nc;
  double* tot_s = dsm + 8 * nc;
  int* cls_s = reinterpret_cast<int*>(tot_s + S);
  const int* lab = ln.labels + ln.lab_off[b];
  const double* __restrict__ tot = a.tot + ((size_t)ln.st_off[b]) * kStatSlices;
  for (int s = tid; s < S; s += TILE_THREADS) {
    cls_s[s] = state_class(lab, s, raw);
    double acc = 0.0;
#pragma unroll
    for (int z = 0; z < kStatSlices; z++) acc += tot[(size_t)z * S + s];   // fixed order: deterministic
    tot_s[s] = fmax(1e-9, acc);
  }
  __syncthreads();
  const float* __restrict__ ep_l = a.lmatch + ln.lat_off[b];
  for (int c = warp; c < ncols; c += TILE_THREADS / 32) {
    const int t = t0 + c;
    const float* ep = ep_l + (size_t)t * S;
    for (int k = lane; k < nc; k += 32) acc[k] = 0.0;
    __syncwarp();
    // epath(t,s) /= total_s (Float /= double, ctc.cc:88), then project onto classes with double accumulators (:96-103)
    if (raw) {
      for (int s = lane; s < S; s += 32) atomicAdd(&acc[cls_s[s]], (double)(float)((double)ep[s] / tot_s[s]));
      __syncwarp();
    } else {
      double blank = 0.0;                                  // all even states are class 0
      for (int s = 2 * lane; s < S; s += 64) blank += (double)(float)((double)ep[s] / tot_s[s]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) blank += __shfl_xor_sync(0xffffffffu, blank, o);
      for (int s = 2 * lane + 1; s < S; s += 64) atomicAdd(&acc[cls_s[s]], (double)(float)((double)ep[s] / tot_s[s]));
      __syncwarp();
      if (lane == 0) acc[0] += blank;
      __syncwarp();
    }
    double rt = 0.0;                                       // row total of the Float-rounded values  ctc.cc:104-109
    for (int k = lane; k < nc; k += 32) rt += (double)(float)acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rt += __shfl_xor_sync(0xffffffffu, rt, o);
    rt = fmax(rt, 1e-9);
    const size_t row = (size_t)(off + t) * nc;
    float mv = -INFINITY;
    int mi = -1;
    for (int k = lane; k < nc; k += 32) {
      const float v = (float)((double)(float)acc[k] / rt);
      a.aligned[row + k] = v;
      a.delta[row + k] = v - a.out[row + k];               // outputs[t].d = aligned - outputs  clstmhl.h:211-212
      if (!(v < mv)) { mv = v; mi = k; }                   // argmax, ties -> last index (tensor.h:357-366)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov > mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; }
    }
    if (lane == 0) { a.amax[off + t] = mi; a.amaxv[off + t] = mv; }
    __syncwarp();
  }
}

}  // namespace

int ctc_align(cudaStream_t st, const Lines& ln, const CtcArgs& a) {
  const int ncp = a.nc | 1;
  const size_t sm_a = (size_t)(TC * ncp + 2 * TC) * sizeof(float) + (size_t)kCtcMaxStates * sizeof(int);
  const size_t sm_d = (size_t)(8 * a.nc + kCtcMaxStates) * sizeof(double) + (size_t)kCtcMaxStates * sizeof(int);
  ctc_lmatch_kernel<<<ln.ntiles, TILE_THREADS, sm_a, st>>>(ln, a);
  ctc_lattice_kernel<<<ln.B, 64 * kLatWarps, 0, st>>>(ln, a);
  ctc_max_kernel<<<dim3(ln.B, kStatSlices), 256, 0, st>>>(ln, a);
  ctc_sum_kernel<<<dim3(ln.B, kStatSlices), 256, 0, st>>>(ln, a);
  ctc_posterior_kernel<<<ln.ntiles, TILE_THREADS, sm_d, st>>>(ln, a);
  return 5;
}

int ctc_configure() {
  // worst case dynamic smem: nc = kCtcMaxClasses
  const size_t sm_a = (size_t)(TC * (kCtcMaxClasses | 1) + 2 * TC) * sizeof(float) + (size_t)kCtcMaxStates * sizeof(int);
  const size_t sm_d = (size_t)(8 * kCtcMaxClasses + kCtcMaxStates) * sizeof(double) + (size_t)kCtcMaxStates * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(ctc_lmatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(ctc_posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_d);
  return (int)e;
}

}  // namespace cb200
