// normalizer_oracle.cc -- CPU restatement of the reference's text-line normalizers (TEST INFRASTRUCTURE, like
// clstm_oracle.cc: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it).
//
// Restates /root/reference/extras.cc:
//   gauss1d            :57-87     FIR Gaussian, mask to 3 sigma, edges clamped, double accumulation of float products
//   gauss2d            :111-123   first along y (inside a column, sigma sy), then along x (sigma sx)
//   bilin              :133-145   bilinear sample, indices clamped, mixed float/double arithmetic as written there
//   MeanNormalizer     :154-198   intensity-weighted mean row and mean absolute deviation, uniform scale
//   argmax1            :200-212   per column, ties -> LAST row
//   add_smear          :214-225   exponentially smeared copy of the line, *1e-3, "to avoid singularities"
//   CenterNormalizer   :227-285   smoothed argmax centre line, r = int(range*mad+1), shear-free resampling around it
// Images are Tensor2 image(i, j), i = x (column, 0..w-1), j = y (row, 0..h-1), stored i + j*w (Eigen col-major).
// Parity status: UNPINNED upstream -- the reference holds no golden vectors or tests for the normalizers and cannot
// be built here (Eigen, libpng absent), so this file is pinned only by properties (tests/test_oracle_normalizer.py).
// Arithmetic types follow the reference expression by expression; built with -ffp-contract=off like the reference.
// One interpretation is recorded here because it changes roundings: `fabs(j - center(i))` in CenterNormalizer::measure
// has a float argument; with libstdc++'s <math.h> the float overload is selected, so the product stays in float.
#include <cmath>
#include <cstring>
#include <vector>

namespace {

typedef float Float;

struct Img {  // view: (i, j) -> p[i + j*w]
  Float* p; int w, h;
  Float& operator()(int i, int j) const { return p[(size_t)i + (size_t)j * w]; }
};

std::vector<float> gauss_mask(float sigma, int& range) {  // extras.cc:60-69
  range = 1 + int(3.0 * sigma);
  std::vector<float> mask(2 * range + 1);
  for (int i = 0; i <= range; i++) {
    double y = exp(-i * i / 2.0 / sigma / sigma);
    mask[range + i] = mask[range - i] = y;
  }
  float total = 0.0;
  for (size_t i = 0; i < mask.size(); i++) total += mask[i];
  for (size_t i = 0; i < mask.size(); i++) mask[i] /= total;
  return mask;
}

void gauss1d(std::vector<float>& out, const std::vector<float>& in, float sigma) {  // extras.cc:57-87
  int range;
  std::vector<float> mask = gauss_mask(sigma, range);
  const int n = (int)in.size(), m = (int)mask.size();
  out.resize(n);
  for (int i = 0; i < n; i++) {
    double total = 0.0;
    for (int j = 0; j < m; j++) {
      int index = i + j - range;
      if (index < 0) index = 0;
      if (index >= n) index = n - 1;
      total += in[index] * mask[j];
    }
    out[i] = Float(total);
  }
}

void gauss2d(Img a, float sx, float sy) {  // extras.cc:111-123
  std::vector<float> r, s;
  for (int i = 0; i < a.w; i++) {
    r.resize(a.h);
    for (int j = 0; j < a.h; j++) r[j] = a(i, j);
    gauss1d(s, r, sy);
    for (int j = 0; j < a.h; j++) a(i, j) = s[j];
  }
  for (int j = 0; j < a.h; j++) {
    r.resize(a.w);
    for (int i = 0; i < a.w; i++) r[i] = a(i, j);
    gauss1d(s, r, sx);
    for (int i = 0; i < a.w; i++) a(i, j) = s[i];
  }
}

inline int clipi(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

inline Float bilin(Img a, float x, float y) {  // extras.cc:133-145
  int w = a.w, h = a.h;
  int i = (int)floor(x);
  int j = (int)floor(y);
  float l = x - i;
  float m = y - j;
  float s00 = a(clipi(i, w), clipi(j, h));
  float s01 = a(clipi(i, w), clipi(j + 1, h));
  float s10 = a(clipi(i + 1, w), clipi(j, h));
  float s11 = a(clipi(i + 1, w), clipi(j + 1, h));
  return ((1.0 - l) * ((1.0 - m) * s00 + m * s01) + l * ((1.0 - m) * s10 + m * s11));
}

}  // namespace

extern "C" {

// Gaussian mask exactly as the reference builds it (used to check the product's host-side mask builder).
int oracle_gauss_mask(float sigma, float* mask, int cap) {
  int range;
  std::vector<float> m = gauss_mask(sigma, range);
  if ((int)m.size() > cap) return -(int)m.size();
  memcpy(mask, m.data(), m.size() * sizeof(float));
  return range;
}

// CenterNormalizer::measure (extras.cc:237-256).  center: w floats.  Returns r.  smooth_out (optional): w*h floats.
float oracle_center_measure(const float* line_, int w, int h, float range, float smooth2d, float smooth1d,
                            float* center_out, float* smooth_out) {
  std::vector<float> linev(line_, line_ + (size_t)w * h), smoothv(linev);
  Img line{linev.data(), w, h}, smooth{smoothv.data(), w, h};
  gauss2d(smooth, h * smooth2d, h * 0.5);
  for (int j = 0; j < h; j++) {  // add_smear, extras.cc:214-225
    double v = 0.0;
    for (int i = 0; i < w; i++) {
      v = v * 0.9 + line(i, j);
      smooth(i, j) += fmin(1.0, v) * 1e-3;
    }
  }
  std::vector<float> a(w), center;
  for (int i = 0; i < w; i++) {  // argmax1, extras.cc:200-212
    float mv = smooth(i, 0);
    float mj = 0;
    for (int j = 1; j < h; j++) {
      if (smooth(i, j) < mv) continue;
      mv = smooth(i, j);
      mj = j;
    }
    a[i] = mj;
  }
  gauss1d(center, a, h * smooth1d);
  float s1 = 0.0;
  float sy = 0.0;
  for (int i = 0; i < w; i++) {
    for (int j = 0; j < h; j++) {
      s1 += line(i, j);
      sy += line(i, j) * std::fabs(j - center[i]);
    }
  }
  float mad = sy / s1;
  float r = int(range * mad + 1);
  memcpy(center_out, center.data(), w * sizeof(float));
  if (smooth_out) memcpy(smooth_out, smoothv.data(), (size_t)w * h * sizeof(float));
  return r;
}

// target width of CenterNormalizer::normalize (extras.cc:275-276)
int oracle_center_width(int w, float r, int target_height) {
  float scale = (2.0 * r) / target_height;
  int tw = int(w / scale);
  return tw > 1 ? tw : 1;
}

// CenterNormalizer::normalize (extras.cc:272-284).  out: target_width x target_height, (i, j) at i + j*target_width.
void oracle_center_normalize(const float* in_, int w, int h, const float* center, float r, int target_height, float* out) {
  Img in{const_cast<float*>(in_), w, h};
  float scale = (2.0 * r) / target_height;
  int target_width = oracle_center_width(w, r, target_height);
  for (int i = 0; i < target_width; i++) {
    for (int j = 0; j < target_height; j++) {
      float x = scale * i;
      float y = scale * (j - target_height / 2) + center[int(x)];
      out[(size_t)i + (size_t)j * target_width] = bilin(in, x, y);
    }
  }
}

// MeanNormalizer::measure (extras.cc:162-183): y_mean, y_mad in double
void oracle_mean_measure(const float* line_, int w, int h, double* y_mean, double* y_mad) {
  Img line{const_cast<float*>(line_), w, h};
  double sy = 0, s1 = 0;
  for (int i = 0; i < w; i++)
    for (int j = 0; j < h; j++) {
      sy += line(i, j) * j;
      s1 += line(i, j);
    }
  *y_mean = sy / s1;
  sy = 0; s1 = 0;
  for (int i = 0; i < w; i++)
    for (int j = 0; j < h; j++) {
      sy += line(i, j) * fabs(j - *y_mean);
      s1 += line(i, j);
    }
  *y_mad = sy / s1;
}
int oracle_mean_width(int w, double y_mad, float vscale, float range, int target_height) {  // extras.cc:185-188
  float actual = vscale * 2 * range * y_mad;
  float scale = actual / target_height;
  return int(w / scale);
}
void oracle_mean_normalize(const float* in_, int w, int h, double y_mean, double y_mad, float vscale, float range,
                           int target_height, float* out) {  // extras.cc:184-197
  Img in{const_cast<float*>(in_), w, h};
  float actual = vscale * 2 * range * y_mad;
  float scale = actual / target_height;
  int nw = int(w / scale);
  int nh = target_height;
  for (int i = 0; i < nw; i++)
    for (int j = 0; j < nh; j++) {
      float x = scale * i;
      float y = scale * (j - target_height / 2) + y_mean;
      out[(size_t)i + (size_t)j * nw] = bilin(in, x, y);
    }
}

}  // extern "C"
