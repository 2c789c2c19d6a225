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

// ---- separable Gaussian (gauss1d semantics: total is double, the product is float, taps in increasing order).
// Each thread produces FOUR consecutive outputs with a sliding window: source element p is loaded once and feeds tap
// k = p - e of output e (e = 0..3), the four mask values slide through registers, so a tap costs FMUL + F2F + DADD plus
// a quarter of the loads.  Every output still accumulates its taps in the reference's order.  Taps outside [0, 2R] are
// skipped with predicates in the peeled first / last three positions (never by adding zeros).
constexpr int GT = 128;                       // threads per block
__device__ __forceinline__ void tap(double& total, float s, float m) { total = __dadd_rn(total, (double)__fmul_rn(s, m)); }

// along x: block = one row j, 4*GT consecutive outputs; the clamped source segment and the mask are staged in smem
__global__ void __launch_bounds__(GT) norm_gauss_x_kernel(NormLines nl, const float* __restrict__ src,
                                                          float* __restrict__ dst) {
  extern __shared__ float gsm[];
  const int b = blockIdx.z, j = blockIdx.y;
  const int w = nl.W[b], h = nl.H[b];
  const int i0 = blockIdx.x * (4 * GT);
  if (j >= h || i0 >= w) return;
  const int range = nl.mrange[3 * b + 1], m = 2 * range + 1;
  float* mask_s = gsm;                        // [m]
  float* seg = gsm + ((m + 3) & ~3);          // [4*GT + 2*range]: seg[q] = row[clamp(i0 + q - range)]
  const float* __restrict__ mask = nl.masks + nl.moff[3 * b + 1];
  const float* __restrict__ row = src + nl.poff[b] + (size_t)j * w;
  for (int k = threadIdx.x; k < m; k += GT) mask_s[k] = mask[k];
  for (int q = threadIdx.x; q < 4 * GT + 2 * range; q += GT) seg[q] = row[clampi(i0 + q - range, w)];
  __syncthreads();
  const float* __restrict__ sp = seg + 4 * threadIdx.x;     // sp[p] feeds tap p - e of output e
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  // p = 0, 1, 2: outputs e <= p only
  { const float s0 = sp[0]; tap(t0, s0, mask_s[0]); }
  { const float s1 = sp[1]; if (1 <= 2 * range) tap(t0, s1, mask_s[1]); tap(t1, s1, mask_s[0]); }
  { const float s2 = sp[2]; if (2 <= 2 * range) tap(t0, s2, mask_s[2]); if (1 <= 2 * range) tap(t1, s2, mask_s[1]); tap(t2, s2, mask_s[0]); }
  float m1 = (2 <= 2 * range) ? mask_s[2] : 0.f, m2 = (1 <= 2 * range) ? mask_s[1] : 0.f, m3 = mask_s[0];
#pragma unroll 4
  for (int p = 3; p <= 2 * range; p++) {      // all four taps valid
    const float s = sp[p], m0 = mask_s[p];
    tap(t0, s, m0); tap(t1, s, m1); tap(t2, s, m2); tap(t3, s, m3);
    m3 = m2; m2 = m1; m1 = m0;
  }
  // p = 2R+1 .. 2R+3: output e needs k = p - e <= 2R.  (for range >= 2 the window registers hold mask[2R], [2R-1], [2R-2])
  if (2 * range >= 3) {
    { const float s = sp[2 * range + 1]; tap(t1, s, m1); tap(t2, s, m2); tap(t3, s, m3); m3 = m2; m2 = m1; }
    { const float s = sp[2 * range + 2]; tap(t2, s, m2); tap(t3, s, m3); m3 = m2; }
    { const float s = sp[2 * range + 3]; tap(t3, s, m3); }
  } else {                                    // range == 1 (m = 3): finish the short windows explicitly
    for (int e = 1; e < 4; e++) {
      double& t = (e == 1) ? t1 : (e == 2) ? t2 : t3;
      for (int k = max(0, 3 - e); k < m; k++) if (e + k >= 3) tap(t, sp[e + k], mask_s[k]);
    }
  }
  float* __restrict__ out = dst + nl.poff[b] + (size_t)j * w + i0 + 4 * threadIdx.x;
  const int left = w - (i0 + 4 * (int)threadIdx.x);
  if (left > 0) out[0] = (float)t0;
  if (left > 1) out[1] = (float)t1;
  if (left > 2) out[2] = (float)t2;
  if (left > 3) out[3] = (float)t3;
}

// along y: thread = column i, four consecutive rows j0..j0+3; consecutive threads read consecutive columns (coalesced)
__global__ void __launch_bounds__(GT) norm_gauss_y_kernel(NormLines nl, const float* __restrict__ src,
                                                          float* __restrict__ dst) {
  extern __shared__ float gsm[];
  const int b = blockIdx.z;
  const int w = nl.W[b], h = nl.H[b];
  const int j0 = 4 * blockIdx.y;
  if (j0 >= h || (int)(blockIdx.x * GT) >= w) return;
  const int range = nl.mrange[3 * b + 0], m = 2 * range + 1;
  const float* __restrict__ mask = nl.masks + nl.moff[3 * b + 0];
  for (int k = threadIdx.x; k < m; k += GT) gsm[k] = mask[k];
  __syncthreads();
  const int i = blockIdx.x * GT + threadIdx.x;
  if (i >= w) return;
  const float* __restrict__ col = src + nl.poff[b] + i;
  auto at = [&](int p) { return col[(size_t)clampi(j0 + p - range, h) * w]; };   // element feeding tap p - e of output e
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  { const float s0 = at(0); tap(t0, s0, gsm[0]); }
  { const float s1 = at(1); if (1 <= 2 * range) tap(t0, s1, gsm[1]); tap(t1, s1, gsm[0]); }
  { const float s2 = at(2); if (2 <= 2 * range) tap(t0, s2, gsm[2]); if (1 <= 2 * range) tap(t1, s2, gsm[1]); tap(t2, s2, gsm[0]); }
  float m1 = (2 <= 2 * range) ? gsm[2] : 0.f, m2 = (1 <= 2 * range) ? gsm[1] : 0.f, m3 = gsm[0];
#pragma unroll 4
  for (int p = 3; p <= 2 * range; p++) {
    const float s = at(p), m0 = gsm[p];
    tap(t0, s, m0); tap(t1, s, m1); tap(t2, s, m2); tap(t3, s, m3);
    m3 = m2; m2 = m1; m1 = m0;
  }
  if (2 * range >= 3) {
    { const float s = at(2 * range + 1); tap(t1, s, m1); tap(t2, s, m2); tap(t3, s, m3); m3 = m2; m2 = m1; }
    { const float s = at(2 * range + 2); tap(t2, s, m2); tap(t3, s, m3); m3 = m2; }
    { const float s = at(2 * range + 3); tap(t3, s, m3); }
  } else {
    for (int e = 1; e < 4; e++) {
      double& t = (e == 1) ? t1 : (e == 2) ? t2 : t3;
      for (int k = max(0, 3 - e); k < m; k++) if (e + k >= 3) tap(t, at(e + k), gsm[k]);
    }
  }
  float* __restrict__ out = dst + nl.poff[b] + i;
  if (j0 + 0 < h) out[(size_t)(j0 + 0) * w] = (float)t0;
  if (j0 + 1 < h) out[(size_t)(j0 + 1) * w] = (float)t1;
  if (j0 + 2 < h) out[(size_t)(j0 + 2) * w] = (float)t2;
  if (j0 + 3 < h) out[(size_t)(j0 + 3) * w] = (float)t3;
}

// ---- rest of CenterNormalizer::measure: smear -> argmax -> centre line (one block per line), then the sums
constexpr int NT = 256;
constexpr int TILE_I = 32;
__global__ void __launch_bounds__(NT) norm_center_line_kernel(NormLines nl, const float* __restrict__ raw_all,
                                                             float* __restrict__ smooth_all, float* __restrict__ a_all,
                                                             float* __restrict__ center_all) {
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
#pragma unroll 4
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
#pragma unroll 4
    for (int j = 1; j < h; j++) {
      const float s = smooth[i + j * w];
      const bool take = !(s < mv);            // `if (a < mv) continue;` => NaNs and ties take the later row
      mv = take ? s : mv;
      mj = take ? (float)j : mj;
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
}

// the two order-sensitive sums of measure() and r, one block per line
__global__ void __launch_bounds__(NT) norm_center_sums_kernel(NormLines nl, const float* __restrict__ raw_all,
                                                             const float* __restrict__ center_all, float* __restrict__ r_out) {
  extern __shared__ __align__(16) float tile[];   // 32 columns x h rows as a linear stream, see below
  const int b = blockIdx.x, tid = threadIdx.x;
  const int w = nl.W[b], h = nl.H[b];
  const float* __restrict__ raw = raw_all + nl.poff[b];
  const float* __restrict__ center = center_all + nl.coff[b];
  // s1 = sum line(i,j), sy = sum line(i,j) * |j - center(i)|, both float, strictly in (i outer, j inner) order.
  // The block stages TI columns at a time as LINEAR STREAMS in summation order (index di*h + j): `tile` holds the pixels,
  // `prod` the products pixel * |j - center| (computed by all threads, they are independent).  Lane 0 of warp 0 then
  // chains s1 over `tile`, lane 0 of warp 1 chains sy over `prod`, both with 128-bit loads: only the FADD chains are serial.
  const int TI = (h <= 384) ? TILE_I : (h <= 768 ? TILE_I / 2 : TILE_I / 4);   // two streams must fit shared memory
  float* prod = tile + TI * h;
  float s1 = 0.f, sy = 0.f;
  for (int i0 = 0; i0 < w; i0 += TI) {
    const int ni = min(TI, w - i0);
    for (int e = tid; e < h * TI; e += NT) {
      const int j = e / TI, di = e % TI;
      if (di < ni) {
        const float v = raw[i0 + di + j * w];
        tile[di * h + j] = v;
        prod[di * h + j] = __fmul_rn(v, fabsf(__fsub_rn((float)j, center[i0 + di])));
      }
    }
    __syncthreads();
    const int cnt = ni * h;
    if (tid == 0 || tid == 32) {
      const float* __restrict__ seq = tid ? prod : tile;
      float acc = tid ? sy : s1;
      int k = 0;
#pragma unroll 4
      for (; k + 4 <= cnt; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&seq[k]);
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v.x), v.y), v.z), v.w);
      }
      for (; k < cnt; k++) acc = __fadd_rn(acc, seq[k]);
      if (tid) sy = acc; else s1 = acc;
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
  extern __shared__ __align__(16) float tile[];
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
        if (di < ni) tile[di * h + j] = raw[i0 + di + j * w];        // linear stream in summation order
      }
      __syncthreads();
      if (tid == 0) {
        for (int di = 0; di < ni; di++) {
          const float* __restrict__ col = tile + di * h;
          if (pass == 0) {
#pragma unroll 8
            for (int j = 0; j < h; j++) {
              const float v = col[j];
              sy = __dadd_rn(sy, (double)__fmul_rn(v, (float)j));
              s1 = __dadd_rn(s1, (double)v);
            }
          } else {
#pragma unroll 8
            for (int j = 0; j < h; j++) {
              const float v = col[j];
              sy = __dadd_rn(sy, __dmul_rn((double)v, fabs(__dsub_rn((double)j, ym))));
              s1 = __dadd_rn(s1, (double)v);
            }
          }
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

int norm_center_measure(cudaStream_t st, const NormLines& nl, int B, int maxw, int maxh, int maxrange, const float* raw,
                        float* tmp, float* smooth, float* a, float* center, float* r_out) {
  const size_t smem_m = (size_t)((2 * maxrange + 1 + 3) & ~3) * sizeof(float);
  norm_gauss_y_kernel<<<dim3((maxw + GT - 1) / GT, (maxh + 3) / 4, B), GT, smem_m, st>>>(nl, raw, tmp);
  norm_gauss_x_kernel<<<dim3((maxw + 4 * GT - 1) / (4 * GT), maxh, B), GT, smem_m + (size_t)(4 * GT + 2 * maxrange) * sizeof(float), st>>>(nl, tmp, smooth);
  const size_t smem = (size_t)2 * maxh * ((maxh <= 384) ? TILE_I : (maxh <= 768 ? TILE_I / 2 : TILE_I / 4)) * sizeof(float);
  norm_center_line_kernel<<<B, NT, 0, st>>>(nl, raw, smooth, a, center);
  norm_center_sums_kernel<<<B, NT, smem, st>>>(nl, raw, center, r_out);
  return 4;
}
int norm_mean_measure(cudaStream_t st, const NormLines& nl, int B, int maxh, const float* raw, double* ymean, double* ymad) {
  const size_t smem = (size_t)maxh * TILE_I * sizeof(float);
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
  // largest masks: sigma = kNormMaxHeight * smooth2d with smooth2d up to ~1.5 => range <= 1 + 3 * 1.5 * 1024
  constexpr int kMaxRange = 1 + 3 * 3 * kNormMaxHeight / 2;
  cudaError_t e0 = cudaFuncSetAttribute(norm_gauss_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((4 * kMaxRange + 8 + 4 * GT) * sizeof(float)));
  if (e0 == cudaSuccess)
    e0 = cudaFuncSetAttribute(norm_gauss_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * kMaxRange + 8) * sizeof(float)));
  if (e0 != cudaSuccess) return (int)e0;
  cudaError_t e = cudaFuncSetAttribute(norm_center_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 768 * (TILE_I / 2) * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(norm_mean_measure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNormMaxHeight * TILE_I * 4);
  return (int)e;
}

}  // namespace cb200
