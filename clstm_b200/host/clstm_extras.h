// clstm_extras.h -- host-side mirror of the pieces either side of the hot path:
//   INormalizer + make_*Normalizer        extras.h:31-47, extras.cc:146-301   (arithmetic on the DEVICE, normalize.cu)
//   read_png / write_png                  extras.h:49-50, extras.cc:305-560   (own PNG codec on zlib; no libpng here)
//   getenv helpers, print, Trigger, read_text32, read_lines, write_text, basename, levenshtein
//                                          utils.h:40-325, clstm.h:329-351
// Same names, argument meaning and error behaviour as the reference (THROW(const char*)).
#pragma once
#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "clstm_host.h"

namespace ocropus {

// ---- text line normalization -----------------------------------------------------------------------------------
// Images are Tensor2 image(i, j) with i = column (x), j = row (y), ink = 1.
struct INormalizer {  // extras.h:31-43
  int target_height = 48;
  float smooth2d = 1.0;
  float smooth1d = 0.3;
  float range = 4.0;
  float vscale = 1.0;
  virtual ~INormalizer() {}
  virtual void getparams(bool verbose = false) {}
  virtual void measure(Tensor2& line) = 0;
  virtual void normalize(Tensor2& out, Tensor2& in) = 0;
  virtual int kind() const = 0;          // 0 none, 1 mean, 2 center (the C ABI's numbering)
  virtual const char* name() const = 0;
  // {range, smooth2d, smooth1d, vscale} in the order clstm_b200_normalize_batch takes them
  void abi_params(float p[4]) const { p[0] = range; p[1] = smooth2d; p[2] = smooth1d; p[3] = vscale; }
};
INormalizer* make_Normalizer(const std::string& name);   // "none" | "mean" | "center"; else THROW("unknown normalizer name")
INormalizer* make_NoNormalizer();
INormalizer* make_MeanNormalizer();
INormalizer* make_CenterNormalizer();

// ---- PNG ---------------------------------------------------------------------------------------------------------
// read_png: grey value (r+g+b)/(3*255.0) per pixel, image(x, y), y = 0 at the top (extras.cc:537-551); accepts every
// colour type / bit depth libpng's STRIP_16|STRIP_ALPHA|PACKING|EXPAND transforms accept, interlaced files included.
void read_png(Tensor2& image, const char* name);
void write_png(const char* name, Tensor2& image);        // 8-bit RGB, value = floor(clip(v*256, 0, 255.999999))

// ---- utils.h -----------------------------------------------------------------------------------------------------
double now();
std::string basename(std::string s);                     // strips directory AND everything from the first '.' (utils.h:59-71)
std::string read_text(std::string fname, int maxsize = 65536);
std::wstring read_text32(std::string fname, int maxsize = 65536);
void read_lines(std::vector<std::string>& lines, std::string fname);
void write_text(const std::string fname, const std::wstring& data);
void write_text(const std::string fname, const std::string& data);

bool reported_params(const char* name);
void report_param_value(const char* name, const std::string& value);
const char* getsenv(const char* name, const char* dflt);
int getienv(const char* name, int dflt = 0);
double getdenv(const char* name, double dflt = 0);
double getrenv(const char* name, double dflt = 0, bool logscale = true);   // var=7.3 or var=2,8 (log-uniform)
double getuenv(const char* name, double dflt = 0);                          // var=2,8 uniform

inline void print() { std::cout << std::endl; }
inline std::ostream& operator<<(std::ostream& stream, const std::wstring& arg) {   // utils.h:117-120 (writes to cout)
  std::cout << utf32_to_utf8(arg);
  return stream;
}
template <class T>
inline void print(const T& arg) { std::cout << arg << std::endl; }
template <class T, typename... Args>
inline void print(T arg, Args... args) {
  std::cout << arg << " ";
  print(args...);
}

struct Trigger {  // utils.h:274-323 "report every ..." logic
  bool finished = false;
  bool enabled = true;
  int count = 0;
  int every = 1;
  int upto = 0;
  int next = 0;
  int last_trigger = 0;
  int current_trigger = 0;
  Trigger(int every, int upto = -1, int start = 0) : count(start), every(every), upto(upto) {}
  Trigger& skip0() { next += every; return *this; }
  Trigger& enable(bool flag) { enabled = flag; return *this; }
  void rotate() { last_trigger = current_trigger; current_trigger = count; }
  int since() { return count - last_trigger; }
  bool check();
  bool operator()(int current);
  bool operator+=(int incr) { return operator()(count + incr); }
  bool operator++() { return operator()(count + 1); }
};

template <class A, class B>
double levenshtein(A& a, B& b) {  // clstm.h:329-351: unit-cost edit distance, two rolling rows
  const int n = (int)a.size(), m = (int)b.size();
  if (n > m) return levenshtein(b, a);
  std::vector<double> cur(n + 1), prev(n + 1);
  for (int k = 0; k <= n; k++) cur[k] = k;
  for (int i = 1; i <= m; i++) {
    prev.swap(cur);
    cur[0] = i;
    for (int j = 1; j <= n; j++) {
      const double sub = prev[j - 1] + (a[j - 1] != b[i - 1] ? 1 : 0);
      cur[j] = std::fmin(std::fmin(prev[j] + 1, cur[j - 1] + 1), sub);
    }
  }
  return cur[n];
}

void network_info(Network net, std::string prefix);      // clstm.cc:269-277

}  // namespace ocropus
