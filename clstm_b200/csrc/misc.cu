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

// Full<F> output layers (clstm.cc:357-389): y = f(z) element-wise, kind 1 SIG, 2 LIN, 3 TANH, 4 RELU, plus the per-column
// argmax like softmax_rows.  sigmoid is Eigen's unclamped 1/(1+exp(-x)) (clstm_compute.cc:116-118).
__device__ __forceinline__ float full_act(int kind, float x) {
  switch (kind) {
    case 1: return 1.0f / (1.0f + expf(-x));
    case 3: return tanhf(x);
    case 4: return fmaxf(x, 0.f);
    default: return x;
  }
}
__global__ void full_rows_kernel(float* __restrict__ z, int N, int nc, int kind, int* __restrict__ amax,
                                 float* __restrict__ amaxv) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < N; n += nwarps) {
    float* row = z + (size_t)n * nc;
    float mv = -INFINITY;
    int mi = -1;
    for (int c = lane; c < nc; c += 32) {
      const float o = full_act(kind, row[c]);
      row[c] = o;
      if (!(o < mv)) { mv = o; mi = c; }
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
// backward_nonlin0 (clstm_compute.cc:231-267): delta <- f'(y) * delta, f' expressed through the output y
__global__ void full_backward_kernel(float* __restrict__ delta, const float* __restrict__ y, size_t n, int kind) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = y[i], d = delta[i];
    float r;
    switch (kind) {
      case 1: r = v * (1.f - v) * d; break;
      case 3: r = (1.f - v * v) * d; break;
      case 4: r = d * (v > 0.f ? 1.f : 0.f); break;
      default: r = d;
    }
    delta[i] = r;
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
// ------------------------------------------------------------------------------------------------------------------
// share_deltas (clstm.cc:731-744) + sgd_update (clstm.cc:201-217) in ONE kernel over NVLink peer memory.
// Every rank's step derivatives g live in a peer-mapped "comm buffer" [header 1 KiB | g].  Each rank reads ALL ranks'
// g directly over NVLink/NVSwitch (one-shot all-reduce: world-1 remote reads of P floats, 544 KB..6 MB, latency bound,
// no intermediate buffer), adds them in rank order (identical result on every rank), folds the sum into Params.d and
// applies clip + update.  Flag protocol in the header (32-bit epochs, one slot per writer rank):
//   arrive[r] : rank r's g for this epoch is complete          (written remotely by r into every rank's header)
//   depart[r] : rank r has finished reading everybody's g       (ditto) -> the owner may zero / overwrite its g
// Spins are bounded (~8 s) and trap, so a lost peer becomes an error instead of a hang.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned epoch) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) != epoch) {
    if (clock64() - t0 > 16000000000ll) __trap();
  }
}

__device__ __forceinline__ unsigned long long global_ns_() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256) peer_allreduce_update_kernel(PeerArgs a) {
  unsigned* hdr = reinterpret_cast<unsigned*>(a.comm[a.rank]);
  const bool rec = a.stats != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t_start = rec ? global_ns_() : 0;
  // ---- all ranks' g complete?
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(a.comm[threadIdx.x]) + kPeerArrive + a.rank, a.epoch);
  }
  if (threadIdx.x < a.world) spin_until(hdr + kPeerArrive + threadIdx.x, a.epoch);
  __syncthreads();
  const unsigned long long t_data = rec ? global_ns_() : 0;
  // ---- sum over ranks in rank order, fold, clip, update (clstm_compute.cc:553-563)
  const size_t n4 = a.n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    // all peers' loads of this element in flight before the first add (NVLink latency ~2 us: memory-level parallelism is
    // what buys bandwidth here); summation stays in rank order => bit-identical on every rank
    float4 gq[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; r++)
      if (r < a.world) gq[r] = reinterpret_cast<const float4*>(a.comm[r] + kPeerHeaderFloats)[i];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kMaxPeers; r++)
      if (r < a.world) { s.x += gq[r].x; s.y += gq[r].y; s.z += gq[r].z; s.w += gq[r].w; }
    float4 d = reinterpret_cast<float4*>(a.d)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float dd[4] = {d.x + s.x, d.y + s.y, d.z + s.z, d.w + s.w};
    float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (a.clip < 1e6f) { dd[e] = fminf(dd[e], a.clip); dd[e] = fmaxf(dd[e], -a.clip); }
      vv[e] = __fadd_rn(vv[e], __fmul_rn(dd[e], a.lr));
      dd[e] = dd[e] * a.mom;
    }
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(a.d)[i] = make_float4(dd[0], dd[1], dd[2], dd[3]);
  }
  if (blockIdx.x == 0)   // tail (n % 4) elements
    for (size_t i = n4 * 4 + threadIdx.x; i < a.n; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < a.world; r++) s += (a.comm[r] + kPeerHeaderFloats)[i];
      float di = a.d[i] + s;
      if (a.clip < 1e6f) { di = fminf(di, a.clip); di = fmaxf(di, -a.clip); }
      a.v[i] = __fadd_rn(a.v[i], __fmul_rn(di, a.lr));
      a.d[i] = di * a.mom;
    }
  if (rec) {   // time spent waiting for the slowest rank vs. time in the NVLink read + update loop (this block's share)
    const unsigned long long t_end = global_ns_();
    atomicAdd(a.stats + 0, t_data - t_start);
    atomicAdd(a.stats + 1, t_end - t_data);
    atomicAdd(a.stats + 2, 1ull);
  }
  // ---- last block of this rank announces "done reading" to every rank
  __threadfence();
  __syncthreads();
  __shared__ unsigned last_s;
  if (threadIdx.x == 0) last_s = (atomicAdd(hdr + kPeerCounter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last_s && threadIdx.x < a.world) {
    if (threadIdx.x == 0) hdr[kPeerCounter] = 0;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(a.comm[threadIdx.x]) + kPeerDepart + a.rank, a.epoch);
  }
}
// zero the local g once every rank has finished reading it
__global__ void __launch_bounds__(256) peer_zero_kernel(PeerArgs a) {
  const unsigned* hdr = reinterpret_cast<const unsigned*>(a.comm[a.rank]);
  if (threadIdx.x < a.world) spin_until(hdr + kPeerDepart + threadIdx.x, a.epoch);
  __syncthreads();
  float* g = a.comm[a.rank] + kPeerHeaderFloats;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) g[i] = 0.f;
}

}  // namespace

void peer_allreduce_update(cudaStream_t st, const PeerArgs& a) {
  // enough CTAs to keep ~1 MB of peer loads in flight for large nets, never more than one float4 per thread needs
  int blocks = (int)((a.n / 4 + 255) / 256);
  blocks = blocks < 64 ? 64 : (blocks > 592 ? 592 : blocks);
  peer_allreduce_update_kernel<<<blocks, 256, 0, st>>>(a);
  peer_zero_kernel<<<32, 256, 0, st>>>(a);
}

void softmax_rows(cudaStream_t st, float* z, int N, int nc, int* amax, float* amaxv) {
  if (N <= 0) return;
  const int threads = 256;
  int blocks = (N + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  softmax_rows_kernel<<<blocks, threads, 0, st>>>(z, N, nc, amax, amaxv);
}

void full_rows(cudaStream_t st, float* z, int N, int nc, int kind, int* amax, float* amaxv) {
  if (N <= 0) return;
  int blocks = (N + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  full_rows_kernel<<<blocks, 256, 0, st>>>(z, N, nc, kind, amax, amaxv);
}
void full_backward(cudaStream_t st, float* delta, const float* y, size_t n, int kind) {
  if (n == 0 || kind == 2) return;     // LinearLayer: identity
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  full_backward_kernel<<<(int)blocks, 256, 0, st>>>(delta, y, n, kind);
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
