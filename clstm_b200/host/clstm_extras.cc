// clstm_extras.cc -- see clstm_extras.h.  Normalizer arithmetic runs on the device behind clstm_b200_normalize_batch;
// the PNG codec is written against the PNG specification on top of zlib (the reference links libpng, which this
// image does not have); everything else is small host plumbing with the reference's names and behaviour.
#include "clstm_extras.h"

#include <glob.h>
#include <sys/time.h>
#include <zlib.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>

namespace ocropus {
using std::string;
using std::vector;
using std::wstring;

namespace {
void check_abi(int rc) {
  if (rc != 0) throwf("%s", clstm_b200_last_error());
}
}  // namespace

// ------------------------------------------------------------------------------------------------ normalizers
namespace {
// Stand-alone normalizers need a device context: a minimal private handle whose input batch receives the result.
struct DeviceNormalizer : INormalizer {
  clstm_b200_net* h = nullptr;
  int h_height = -1;
  Tensor2 measured;
  bool have = false;
  ~DeviceNormalizer() override { clstm_b200_destroy(h); }
  void ensure() {
    if (h && h_height == target_height) return;
    clstm_b200_destroy(h);
    h = nullptr;
    clstm_b200_cfg cfg;
    cfg.ninput = target_height; cfg.nhidden = 4; cfg.nclasses = 2; cfg.device = 0;
    check_abi(clstm_b200_create(&cfg, &h));
    h_height = target_height;
  }
  void measure(Tensor2& line) override {   // the measurement itself happens on the device together with normalize()
    measured = line;
    have = true;
  }
  void normalize(Tensor2& out, Tensor2& in) override {
    if (kind() != 0) {
      if (!have || in.dimension(0) != measured.dimension(0)) THROW("measure doesn't match normalize");   // extras.cc:274
      if (in.dimension(1) != measured.dimension(1) || in.data != measured.data)
        THROW("normalize() of an image other than the measured one is not supported on the device");
    }
    ensure();
    int W = in.dimension(0), H = in.dimension(1), T = 0;
    float p[4];
    abi_params(p);
    check_abi(clstm_b200_normalize_batch(h, in.ptr(), &W, &H, 1, kind(), p, nullptr, nullptr, &T));
    vector<float> x((size_t)T * target_height);
    check_abi(clstm_b200_get_inputs(h, x.data()));
    out.resize(T, target_height);
    for (int t = 0; t < T; t++)
      for (int j = 0; j < target_height; j++) out(t, j) = x[(size_t)t * target_height + j];
  }
};
struct NoNormalizer : DeviceNormalizer {      // extras.cc:146-152
  int kind() const override { return 0; }
  const char* name() const override { return "none"; }
};
struct MeanNormalizer : DeviceNormalizer {    // extras.cc:154-198
  MeanNormalizer() { range = 1.0; }
  void getparams(bool verbose) override {
    vscale = getrenv("norm_vscale", 1.0);
    range = getrenv("norm_range", 1.0);
    if (verbose) print("mean_normalizer", range, vscale);
  }
  int kind() const override { return 1; }
  const char* name() const override { return "mean"; }
};
struct CenterNormalizer : DeviceNormalizer {  // extras.cc:227-285
  void getparams(bool verbose) override {
    range = getrenv("norm_range", 4.0);
    smooth2d = getrenv("norm_smooth2d", 1.0);
    smooth1d = getrenv("norm_smooth1d", 0.3);
    if (verbose) print("center_normalizer", range, smooth2d, smooth1d);
  }
  int kind() const override { return 2; }
  const char* name() const override { return "center"; }
};
}  // namespace

INormalizer* make_NoNormalizer() { return new NoNormalizer(); }
INormalizer* make_MeanNormalizer() { return new MeanNormalizer(); }
INormalizer* make_CenterNormalizer() { return new CenterNormalizer(); }
INormalizer* make_Normalizer(const string& name) {  // extras.cc:293-300
  if (name == "none") return make_NoNormalizer();
  if (name == "mean") return make_MeanNormalizer();
  if (name == "center") return make_CenterNormalizer();
  THROW("unknown normalizer name");
}

// ------------------------------------------------------------------------------------------------ PNG
namespace {
typedef unsigned char u8;
unsigned be32(const u8* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }
void put32(vector<u8>& v, unsigned x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// undo the scanline filters of one (sub)image in place; `data` holds rows of 1 + stride bytes
void unfilter(u8* data, int rows, int stride, int bpp) {
  vector<u8> zero(stride, 0);
  const u8* prev = zero.data();
  for (int y = 0; y < rows; y++) {
    u8* line = data + (size_t)y * (stride + 1);
    const int ft = line[0];
    u8* cur = line + 1;
    for (int x = 0; x < stride; x++) {
      const int a = x >= bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= bpp ? prev[x - bpp] : 0;
      int v = cur[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: THROW("internal png error");
      }
      cur[x] = (u8)v;
    }
    prev = cur;
  }
}

struct PngImage { int w = 0, h = 0; vector<u8> rgb; };   // always 3 x 8 bit, like the reference's Tensor<uchar,3>

void decode_png(PngImage& img, const vector<u8>& file) {
  static const u8 sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) THROW("internal png error");
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  vector<u8> idat, plte;
  bool have_hdr = false, done = false;
  while (!done && pos + 12 <= file.size()) {
    const unsigned len = be32(&file[pos]);
    const char* type = (const char*)&file[pos + 4];
    if (pos + 12 + (size_t)len > file.size()) THROW("internal png error");
    const u8* body = &file[pos + 8];
    if (crc32(crc32(0L, Z_NULL, 0), &file[pos + 4], len + 4) != be32(body + len)) THROW("internal png error");
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) THROW("internal png error");
      img.w = (int)be32(body); img.h = (int)be32(body + 4);
      depth = body[8]; ctype = body[9]; interlace = body[12];
      if (body[10] != 0 || body[11] != 0 || interlace > 1) THROW("internal png error");
      have_hdr = true;
    } else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
    else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!memcmp(type, "IEND", 4)) done = true;
    pos += 12 + (size_t)len;
  }
  if (!have_hdr || img.w <= 0 || img.h <= 0) THROW("internal png error");
  if (img.w > (1 << 20) || img.h > (1 << 20) || (long long)img.w * img.h > (1LL << 28)) THROW("internal png error");   // absurd header
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: THROW("internal png error");
  }
  const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                        (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                        ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
  if (!depth_ok) THROW("internal png error");
  if (ctype == 3 && plte.size() < 3) THROW("internal png error");
  const int bits = depth * channels, bpp = std::max(1, bits / 8);
  // pass geometry: one pass for plain files, Adam7 otherwise
  static const int X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1},
                   DY[7] = {8, 8, 8, 4, 4, 2, 2};
  const int npass = interlace ? 7 : 1;
  size_t need = 0;
  int pw[7], ph[7];
  for (int p = 0; p < npass; p++) {
    pw[p] = interlace ? (img.w - X0[p] + DX[p] - 1) / DX[p] : img.w;
    ph[p] = interlace ? (img.h - Y0[p] + DY[p] - 1) / DY[p] : img.h;
    if (pw[p] > 0 && ph[p] > 0) need += (size_t)ph[p] * (1 + ((size_t)pw[p] * bits + 7) / 8);
  }
  vector<u8> raw(need);
  uLongf got = (uLongf)need;
  const int zr = uncompress(raw.data(), &got, idat.data(), (uLong)idat.size());
  if (zr != Z_OK || got != need) THROW("internal png error");
  img.rgb.assign((size_t)img.w * img.h * 3, 0);
  const int maxv = (1 << std::min(depth, 8)) - 1;
  size_t off = 0;
  for (int p = 0; p < npass; p++) {
    if (pw[p] <= 0 || ph[p] <= 0) continue;
    const int stride = (pw[p] * bits + 7) / 8;
    unfilter(&raw[off], ph[p], stride, bpp);
    for (int yy = 0; yy < ph[p]; yy++) {
      const u8* row = &raw[off + (size_t)yy * (stride + 1) + 1];
      for (int xx = 0; xx < pw[p]; xx++) {
        int s[4] = {0, 0, 0, 0};
        for (int c = 0; c < channels; c++) {
          const int k = xx * channels + c;
          if (depth == 8) s[c] = row[k];
          else if (depth == 16) s[c] = row[2 * k];                                   // STRIP_16: keep the high byte
          else s[c] = (row[(k * depth) >> 3] >> (8 - depth - ((k * depth) & 7))) & maxv;   // packed, MSB first
        }
        int r, g, b;
        if (ctype == 3) {
          if ((size_t)s[0] * 3 + 2 >= plte.size()) THROW("internal png error");
          r = plte[3 * s[0]]; g = plte[3 * s[0] + 1]; b = plte[3 * s[0] + 2];
        } else if (ctype == 0 || ctype == 4) {
          r = g = b = (depth < 8) ? s[0] * 255 / maxv : s[0];                         // EXPAND scales low-depth grey
        } else {
          r = s[0]; g = s[1]; b = s[2];
        }
        const int x = interlace ? X0[p] + xx * DX[p] : xx, y = interlace ? Y0[p] + yy * DY[p] : yy;
        u8* o = &img.rgb[((size_t)y * img.w + x) * 3];
        o[0] = (u8)r; o[1] = (u8)g; o[2] = (u8)b;
      }
    }
    off += (size_t)ph[p] * (stride + 1);
  }
}

void chunk(vector<u8>& out, const char* type, const vector<u8>& body) {
  put32(out, (unsigned)body.size());
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), body.begin(), body.end());
  put32(out, (unsigned)crc32(crc32(0L, Z_NULL, 0), &out[start], (uInt)(out.size() - start)));
}
void encode_png(vector<u8>& out, int w, int h, const vector<u8>& rgb) {   // 8-bit RGB, filter 0, 300x300 pHYs (extras.cc:493-500)
  static const u8 sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  out.assign(sig, sig + 8);
  vector<u8> hdr;
  put32(hdr, w); put32(hdr, h);
  hdr.push_back(8); hdr.push_back(2); hdr.push_back(0); hdr.push_back(0); hdr.push_back(0);
  chunk(out, "IHDR", hdr);
  vector<u8> phys;
  put32(phys, 300); put32(phys, 300); phys.push_back(1);
  chunk(out, "pHYs", phys);
  vector<u8> raw((size_t)h * (1 + 3 * (size_t)w));
  for (int y = 0; y < h; y++) {
    raw[(size_t)y * (1 + 3 * w)] = 0;
    memcpy(&raw[(size_t)y * (1 + 3 * w) + 1], &rgb[(size_t)y * w * 3], (size_t)3 * w);
  }
  uLongf clen = compressBound((uLong)raw.size());
  vector<u8> z(clen);
  if (compress2(z.data(), &clen, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK) THROW("internal png error");
  z.resize(clen);
  chunk(out, "IDAT", z);
  chunk(out, "IEND", vector<u8>());
}
}  // namespace

void read_png(Tensor2& image, const char* name) {
  FILE* stream = fopen(name, "r");
  if (!stream) THROW("error on open");
  vector<u8> file;
  u8 buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), stream)) > 0) file.insert(file.end(), buf, buf + n);
  fclose(stream);
  PngImage img;
  decode_png(img, file);
  image.resize(img.w, img.h);
  for (int i = 0; i < img.w; i++)
    for (int j = 0; j < img.h; j++) {
      const u8* p = &img.rgb[((size_t)j * img.w + i) * 3];
      image(i, j) = (p[0] + p[1] + p[2]) / (3 * 255.0);
    }
}
void write_png(const char* name, Tensor2& image) {
  const int w = image.dimension(0), h = image.dimension(1);
  vector<u8> rgb((size_t)w * h * 3);
  for (int i = 0; i < w; i++)
    for (int j = 0; j < h; j++) {
      double v = image(i, j) * 256;
      v = v < 0.0 ? 0.0 : (v > 255.999999 ? 255.999999 : v);
      const u8 value = (u8)floor(v);
      u8* p = &rgb[((size_t)j * w + i) * 3];
      p[0] = p[1] = p[2] = value;
    }
  vector<u8> out;
  encode_png(out, w, h, rgb);
  FILE* stream = fopen(name, "w");
  if (!stream) THROW("error on open");
  const bool ok = fwrite(out.data(), 1, out.size(), stream) == out.size();
  fclose(stream);
  if (!ok) THROW("internal png error");
}

// ------------------------------------------------------------------------------------------------ utils
double now() {
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}
string basename(string s) {
  const size_t slash = s.rfind('/');
  const size_t start = slash == string::npos ? 0 : slash + 1;
  const size_t dot = s.find('.', start);
  return dot == string::npos ? s : s.substr(0, dot);
}
string read_text(string fname, int maxsize) {
  std::ifstream stream(fname);
  string buf((size_t)maxsize - 1, '\0');
  stream.read(&buf[0], maxsize - 1);
  size_t n = (size_t)stream.gcount();
  while (n > 0 && buf[n - 1] == '\n') n--;
  return buf.substr(0, n);
}
wstring read_text32(string fname, int maxsize) { return utf8_to_utf32(read_text(fname, maxsize)); }
void read_lines(vector<string>& lines, string fname) {
  std::ifstream stream(fname);
  string line;
  lines.clear();
  while (getline(stream, line)) lines.push_back(line);
}
void write_text(const string fname, const wstring& data) { write_text(fname, utf32_to_utf8(data)); }
void write_text(const string fname, const string& data) {
  std::ofstream stream(fname);
  stream << data << std::endl;
}

bool reported_params(const char* name) {   // each variable is reported once
  static std::set<string> seen;
  return !seen.insert(name).second;
}
void report_param_value(const char* name, const string& value) {   // utils.h:160-166
  const char* flag = getenv("params");
  if (flag && !atoi(flag)) return;
  if (reported_params(name)) return;
  std::cerr << "#: " << name << " = " << value << std::endl;
}
namespace {
string num(double x) {
  char b[64];
  snprintf(b, sizeof b, "%g", x);
  return b;
}
}  // namespace
const char* getsenv(const char* name, const char* dflt) {
  const char* result = getenv(name) ? getenv(name) : dflt;
  report_param_value(name, result);
  return result;
}
int getienv(const char* name, int dflt) {
  const int result = getenv(name) ? atoi(getenv(name)) : dflt;
  report_param_value(name, std::to_string(result));
  return result;
}
double getdenv(const char* name, double dflt) {
  const double result = getenv(name) ? atof(getenv(name)) : dflt;
  report_param_value(name, num(result));
  return result;
}
double getrenv(const char* name, double dflt, bool) {
  const char* s = getenv(name);
  if (!s) return dflt;
  float lo, hi;
  if (sscanf(s, "%g,%g", &lo, &hi) == 2) {
    const double x = exp(log(lo) + drand48() * (log(hi) - log(lo)));
    report_param_value(name, num(x));
    return x;
  }
  if (sscanf(s, "%g", &lo) == 1) {
    report_param_value(name, num(lo));
    return lo;
  }
  THROW("bad format for getrenv");
}
double getuenv(const char* name, double dflt) {
  const char* s = getenv(name);
  if (!s) return dflt;
  float lo, hi;
  if (sscanf(s, "%g,%g", &lo, &hi) == 2) {
    const double x = lo + drand48() * (hi - lo);
    report_param_value(name, num(x));
    return x;
  }
  if (sscanf(s, "%g", &lo) == 1) {
    report_param_value(name, num(lo));
    return lo;
  }
  THROW("bad format for getuenv");
}

bool Trigger::check() {
  if (finished) THROW("Trigger: already finished");       // assert(!finished) upstream
  if (upto > 0 && count >= upto - 1) {
    finished = true;
    rotate();
    return true;
  }
  if (every == 0) return false;
  if (count < next) return false;
  while (count >= next) next += every;
  rotate();
  return true;
}
bool Trigger::operator()(int current) {
  if (finished || current < count) THROW("Trigger: bad call sequence");
  count = current;
  return check();
}

void network_info(Network net, string prefix) {
  const string nprefix = prefix + "." + net->kind;
  const Float learning_rate = (double)net->attr.get("learning_rate");
  const Float momentum = (double)net->attr.get("momentum");
  std::cout << nprefix << ": " << learning_rate << " " << momentum << " ";
  std::cout << "in " << net->inputs.size() << " " << net->ninput() << " ";
  std::cout << "out " << net->outputs.size() << " " << net->noutput() << std::endl;
  for (auto s : net->sub) network_info(s, nprefix);
}

}  // namespace ocropus
