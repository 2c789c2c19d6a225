// misc.cu -- the HBM-bound pointwise pieces of the path: softmax normalisation, clip+SGD update, greedy decode.
#include "kernels.h"

namespace cb200 {
namespace {

__device__ __forceinline__ float limexp_(float x) {  // tensor.h:78-82
  if (x < -30.f) return expf(-30.f);
  if (x > 30.f) return expf(30.f);
  return expf(x);
}

// forward_softmax after the linear part (clstm_compute.cc:334-337): z = limexp(z); z /= colsum(z).
// One warp per column, values kept in registers (nc <= 32*16).
__global__ void softmax_rows_kernel(float* __restrict__ z, int N, int nc, int* __restrict__ amax,
                                    float* __restrict__ amaxv) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < N; n += nwarps) {
    float* row = z + (size_t)n * nc;
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int c = lane + 32 * i;
      v[i] = 0.f;
      if (c < nc) { v[i] = limexp_(row[c]); sum += v[i]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float mv = -INFINITY;
    int mi = -1;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int c = lane + 32 * i;
      if (c < nc) {
        const float o = v[i] / sum;
        row[c] = o;
        if (!(o < mv)) { mv = o; mi = c; }                 // per-column argmax, ties -> last index (tensor.h:357-366)
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov > mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; }
    }
    if (lane == 0) { amax[n] = mi; amaxv[n] = mv; }
  }
}

// Folds this step's derivatives g into the accumulator d (Params.d) and applies sgd_update(Network)
// (clstm.cc:201-217): clip_gradient then v += lr*d ; d *= momentum  (clstm_compute.cc:553-563).
__global__ void sgd_update_kernel(float* __restrict__ v, float* __restrict__ d, float* __restrict__ g, size_t n,
                                  float lr, float mom, float clip, int fold_only) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float di = d[i] + g[i];
    g[i] = 0.f;
    if (!fold_only) {
      if (clip < 1e6f) { di = fminf(di, clip); di = fmaxf(di, -clip); }
      v[i] = __fadd_rn(v[i], __fmul_rn(di, lr));   // two roundings like the CPU path (no FMA contraction)
      di = di * mom;
    }
    d[i] = di;
  }
}

// dst[c][r] = src[r][c] for a handful of small weight matrices in one launch (derived layouts, refreshed after
// every parameter change): R^T for the generic recurrent kernel, W1^T and Wx^T as K-contiguous B operands of the
// dH / dx products.
__global__ void transpose_batch_kernel(TransposeJobs jobs) {
  __shared__ float tile[32][33];
  const TransposeJob j = jobs.job[blockIdx.z];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (c0 >= j.cols || r0 >= j.rows) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < j.rows && c < j.cols) ? j.src[(size_t)r * j.cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < j.rows && c < j.cols) j.dst[(size_t)c * j.rows + r] = tile[threadIdx.x][i];
  }
}

// trivial_decode (ctc.cc:159-194), one CTA per line, from per-column argmax (index, value) arrays.
__global__ void decode_kernel(Lines ln, const int* __restrict__ amax, const float* __restrict__ amax_val,
                              int* __restrict__ classes, int* __restrict__ locs, int* __restrict__ counts,
                              int max_per_line) {
  const int b = blockIdx.x;
  const int T = ln.T[b], off = ln.off[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // in parallel: a run = maximal stretch of non-blank argmax frames; the thread that
  // sits on a run's first frame walks the run (runs are a handful of frames), keeps the first strictly largest
  // probability, and emits (class, frame) iff a blank closes the run.  A block scan orders the emitted symbols.
  __shared__ int wsum[8];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    int flag = 0, mc = -1, mt = -1;
    if (t < T && amax[off + t] != 0 && (t == 0 || amax[off + t - 1] == 0)) {
      float mv = 0.f;
      int u = t;
      for (; u < T; u++) {
        const int index = amax[off + u];
        if (index == 0) break;
        const float v = amax_val[off + u];
        if (v > mv) { mv = v; mc = index; mt = u; }
      }
      flag = (u < T && mc != -1) ? 1 : 0;      // "there should be a 0 at the end anyway": unclosed trailing run is dropped
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < nw; w++) { if (w < warp) woff += wsum[w]; tot += wsum[w]; }
    const int idx = base_s + woff + pre;
    if (flag && idx < max_per_line) {
      classes[(size_t)b * max_per_line + idx] = mc;
      locs[(size_t)b * max_per_line + idx] = mt;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base_s;
}
}  // namespace

void softmax_rows(cudaStream_t st, float* z, int N, int nc, int* amax, float* amaxv) {
  if (N <= 0) return;
  const int threads = 256;
  int blocks = (N + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  softmax_rows_kernel<<<blocks, threads, 0, st>>>(z, N, nc, amax, amaxv);
}

void sgd_update(cudaStream_t st, float* v, float* d, float* g, size_t n, float lr, float mom, float clip,
                int fold_only) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sgd_update_kernel<<<blocks, 256, 0, st>>>(v, d, g, n, lr, mom, clip, fold_only);
}

void transpose_batch(cudaStream_t st, const TransposeJobs& jobs) {
  int mr = 0, mc = 0;
  for (int i = 0; i < jobs.n; i++) { mr = max(mr, jobs.job[i].rows); mc = max(mc, jobs.job[i].cols); }
  if (jobs.n <= 0 || mr <= 0 || mc <= 0) return;
  dim3 grid((mc + 31) / 32, (mr + 31) / 32, jobs.n);
  transpose_batch_kernel<<<grid, dim3(32, 8), 0, st>>>(jobs);
}

void decode_lines(cudaStream_t st, const Lines& ln, const int* argmax_idx, const float* argmax_val, int* classes,
                  int* locs, int* counts, int max_per_line) {
  if (ln.B <= 0) return;
  decode_kernel<<<ln.B, 256, 0, st>>>(ln, argmax_idx, argmax_val, classes, locs, counts, max_per_line);
}

}  // namespace cb200
