// normalize.cu -- text-line normalizers on the device: the step in front of the hot path.
// A batch of raw line images (any height, any width) is measured and resampled straight into the packed network input
// x[(off_b + t) * ni + j], so normalised images never exist on the host.
//
// Replaces /root/reference/extras.cc: CenterNormalizer (:227-285, with gauss1d/gauss2d :57-123, add_smear :214-225,
// argmax1 :200-212, bilin :133-145), MeanNormalizer (:154-198), NoNormalizer (:146-152).
// Bit-exactness is a design goal here because the results feed integer decisions (r, the target width, centre-line
// indices): every float / double operation of the reference is issued with the same operand types and in the same
// order (explicit __fmul_rn / __dadd_rn ... so nvcc cannot contract them into FMAs); the Gaussian masks are built
// on the HOST with libm's exp exactly as gauss1d does (capi.cu) because a device exp could differ in the last bit.
// The two order-sensitive scalar sums of measure() (s1, sy over all pixels) are walked sequentially by one lane each,
// fed through shared-memory tiles that the whole block loads coalesced.
// Images: (i, j) = (column x, row y) at p[i + j*w], like the reference's Tensor2.
#include "kernels.h"

namespace cb200 {
namespace {

__device__ __forceinline__ int clampi(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

// ---- separable Gaussian (gauss1d semantics: total is double, the product is float)
// pass 0: along y with mask 0 of the line (raw -> tmp); pass 1: along x with mask 1 (tmp -> smooth)
__global__ void norm_gauss_kernel(NormLines nl, const float* __restrict__ src, float* __restrict__ dst, int pass) {
  const int b = blockIdx.y;
  const int w = nl.W[b], h = nl.H[b];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= w * h) return;
  const int i = p % w, j = p / w;
  const float* __restrict__ s = src + nl.poff[b];
  const float* __restrict__ mask = nl.masks + nl.moff[3 * b + pass];
  const int range = nl.mrange[3 * b + pass], m = 2 * range + 1;
  double total = 0.0;
  if (pass == 0) {
    for (int k = 0; k < m; k++) {
      const int jj = clampi(j + k - range, h);
      total = __dadd_rn(total, (double)__fmul_rn(s[i + jj * w], mask[k]));
    }
  } else {
    const float* __restrict__ row = s + j * w;
    for (int k = 0; k < m; k++) {
      const int ii = clampi(i + k - range, w);
      total = __dadd_rn(total, (double)__fmul_rn(row[ii], mask[k]));
    }
  }
  dst[nl.poff[b] + p] = (float)total;
}

// ---- rest of CenterNormalizer::measure, one block per line
constexpr int NT = 256;
constexpr int TILE_I = 32;
__global__ void __launch_bounds__(NT) norm_center_tail_kernel(NormLines nl, const float* __restrict__ raw_all,
                                                             float* __restrict__ smooth_all, float* __restrict__ a_all,
                                                             float* __restrict__ center_all, float* __restrict__ r_out) {
  extern __shared__ float tile[];                 // [h][TILE_I + 1]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int w = nl.W[b], h = nl.H[b];
  const float* __restrict__ raw = raw_all + nl.poff[b];
  float* __restrict__ smooth = smooth_all + nl.poff[b];
  float* __restrict__ a = a_all + nl.coff[b];
  float* __restrict__ center = center_all + nl.coff[b];
  // add_smear: row j is a sequential recurrence along x
  for (int j = tid; j < h; j += NT) {
    double v = 0.0;
    const float* rr = raw + j * w;
    float* sr = smooth + j * w;
    for (int i = 0; i < w; i++) {
      v = __dadd_rn(__dmul_rn(v, 0.9), (double)rr[i]);
      sr[i] = (float)__dadd_rn((double)sr[i], __dmul_rn(fmin(1.0, v), 1e-3));
    }
  }
  __syncthreads();
  // argmax1: per column, ties -> last row
  for (int i = tid; i < w; i += NT) {
    float mv = smooth[i];
    float mj = 0.f;
    for (int j = 1; j < h; j++) {
      const float s = smooth[i + j * w];
      if (s < mv) continue;
      mv = s;
      mj = (float)j;
    }
    a[i] = mj;
  }
  __syncthreads();
  // centre line = gauss1d(a, h * smooth1d)
  {
    const float* __restrict__ mask = nl.masks + nl.moff[3 * b + 2];
    const int range = nl.mrange[3 * b + 2], m = 2 * range + 1;
    for (int i = tid; i < w; i += NT) {
      double total = 0.0;
      for (int k = 0; k < m; k++) total = __dadd_rn(total, (double)__fmul_rn(a[clampi(i + k - range, w)], mask[k]));
      center[i] = (float)total;
    }
  }
  __syncthreads();
  // s1 = sum line(i,j), sy = sum line(i,j) * |j - center(i)|, both float, strictly in (i outer, j inner) order:
  // lane 0 of warp 0 carries s1, lane 0 of warp 1 carries sy; everybody stages tiles of 32 columns.
  float s1 = 0.f, sy = 0.f;
  for (int i0 = 0; i0 < w; i0 += TILE_I) {
    const int ni = min(TILE_I, w - i0);
    for (int e = tid; e < h * TILE_I; e += NT) {
      const int j = e / TILE_I, di = e % TILE_I;
      tile[j * (TILE_I + 1) + di] = (di < ni) ? raw[i0 + di + j * w] : 0.f;
    }
    __syncthreads();
    if (tid == 0) {
      for (int di = 0; di < ni; di++)
        for (int j = 0; j < h; j++) s1 = __fadd_rn(s1, tile[j * (TILE_I + 1) + di]);
    } else if (tid == 32) {
      for (int di = 0; di < ni; di++) {
        const float c = center[i0 + di];
        for (int j = 0; j < h; j++)
          sy = __fadd_rn(sy, __fmul_rn(tile[j * (TILE_I + 1) + di], fabsf(__fsub_rn((float)j, c))));
      }
    }
    __syncthreads();
  }
  __shared__ float sy_s;
  if (tid == 32) sy_s = sy;
  __syncthreads();
  if (tid == 0) {
    const float mad = __fdiv_rn(sy_s, s1);
    r_out[b] = (float)(int)__fadd_rn(__fmul_rn(nl.range, mad), 1.f);
  }
}

// ---- MeanNormalizer::measure: two sequential double sums per line (extras.cc:162-183)
__global__ void __launch_bounds__(NT) norm_mean_measure_kernel(NormLines nl, const float* __restrict__ raw_all,
                                                              double* __restrict__ ym_out, double* __restrict__ yd_out) {
  extern __shared__ float tile[];
  __shared__ double ymean_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int w = nl.W[b], h = nl.H[b];
  const float* __restrict__ raw = raw_all + nl.poff[b];
  for (int pass = 0; pass < 2; pass++) {
    double sy = 0.0, s1 = 0.0;
    const double ym = pass ? ymean_s : 0.0;
    for (int i0 = 0; i0 < w; i0 += TILE_I) {
      const int ni = min(TILE_I, w - i0);
      for (int e = tid; e < h * TILE_I; e += NT) {
        const int j = e / TILE_I, di = e % TILE_I;
        tile[j * (TILE_I + 1) + di] = (di < ni) ? raw[i0 + di + j * w] : 0.f;
      }
      __syncthreads();
      if (tid == 0) {
        for (int di = 0; di < ni; di++)
          for (int j = 0; j < h; j++) {
            const float v = tile[j * (TILE_I + 1) + di];
            if (pass == 0) sy = __dadd_rn(sy, (double)__fmul_rn(v, (float)j));
            else sy = __dadd_rn(sy, __dmul_rn((double)v, fabs(__dsub_rn((double)j, ym))));
            s1 = __dadd_rn(s1, (double)v);
          }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double q = __ddiv_rn(sy, s1);
      if (pass == 0) { ymean_s = q; ym_out[b] = q; }
      else yd_out[b] = q;
    }
    __syncthreads();
  }
}

// ---- resampling (CenterNormalizer::normalize :272-284, MeanNormalizer::normalize :184-197, NoNormalizer)
__device__ __forceinline__ float bilin(const float* __restrict__ a, int w, int h, float x, float y) {
  const int i = (int)floorf(x), j = (int)floorf(y);
  const float l = __fsub_rn(x, (float)i), m = __fsub_rn(y, (float)j);
  const float s00 = a[clampi(i, w) + clampi(j, h) * w], s01 = a[clampi(i, w) + clampi(j + 1, h) * w];
  const float s10 = a[clampi(i + 1, w) + clampi(j, h) * w], s11 = a[clampi(i + 1, w) + clampi(j + 1, h) * w];
  const double om = __dsub_rn(1.0, (double)m), ol = __dsub_rn(1.0, (double)l);
  const double t0 = __dadd_rn(__dmul_rn(om, (double)s00), (double)__fmul_rn(m, s01));
  const double t1 = __dadd_rn(__dmul_rn(om, (double)s10), (double)__fmul_rn(m, s11));
  return (float)__dadd_rn(__dmul_rn(ol, t0), __dmul_rn((double)l, t1));
}
// one thread per output element; j (feature) fastest so that the packed input rows are written coalesced
__global__ void norm_resample_kernel(NormLines nl, const float* __restrict__ raw_all, const float* __restrict__ center_all,
                                     const float* __restrict__ scale_all, const double* __restrict__ ymean_all,
                                     const int* __restrict__ T, const int* __restrict__ off, float* __restrict__ x, int ni,
                                     int kind) {
  const int b = blockIdx.y;
  const int tw = T[b];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= tw * ni) return;
  const int i = e / ni, j = e % ni;
  const int w = nl.W[b], h = nl.H[b];
  const float* __restrict__ raw = raw_all + nl.poff[b];
  float v;
  if (kind == 0) {
    v = raw[i + j * w];
  } else {
    const float scale = scale_all[b];
    const float xx = __fmul_rn(scale, (float)i);
    float yy;
    if (kind == 2) yy = __fadd_rn(__fmul_rn(scale, (float)(j - ni / 2)), center_all[nl.coff[b] + (int)xx]);
    else yy = (float)__dadd_rn((double)__fmul_rn(scale, (float)(j - ni / 2)), ymean_all[b]);
    v = bilin(raw, w, h, xx, yy);
  }
  x[(size_t)(off[b] + i) * ni + j] = v;
}

}  // namespace

int norm_center_measure(cudaStream_t st, const NormLines& nl, int B, int maxpix, int maxh, const float* raw, float* tmp,
                        float* smooth, float* a, float* center, float* r_out) {
  dim3 grid((maxpix + 255) / 256, B);
  norm_gauss_kernel<<<grid, 256, 0, st>>>(nl, raw, tmp, 0);
  norm_gauss_kernel<<<grid, 256, 0, st>>>(nl, tmp, smooth, 1);
  const size_t smem = (size_t)maxh * (TILE_I + 1) * sizeof(float);
  norm_center_tail_kernel<<<B, NT, smem, st>>>(nl, raw, smooth, a, center, r_out);
  return 3;
}
int norm_mean_measure(cudaStream_t st, const NormLines& nl, int B, int maxh, const float* raw, double* ymean, double* ymad) {
  const size_t smem = (size_t)maxh * (TILE_I + 1) * sizeof(float);
  norm_mean_measure_kernel<<<B, NT, smem, st>>>(nl, raw, ymean, ymad);
  return 1;
}
int norm_resample(cudaStream_t st, const NormLines& nl, int B, int maxT, const float* raw, const float* center,
                  const float* scale, const double* ymean, const int* T, const int* off, float* x, int ni, int kind) {
  dim3 grid((maxT * ni + 255) / 256, B);
  norm_resample_kernel<<<grid, 256, 0, st>>>(nl, raw, center, scale, ymean, T, off, x, ni, kind);
  return 1;
}
int norm_configure() {
  cudaError_t e = cudaFuncSetAttribute(norm_center_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNormMaxHeight * (TILE_I + 1) * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(norm_mean_measure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNormMaxHeight * (TILE_I + 1) * 4);
  return (int)e;
}

}  // namespace cb200
