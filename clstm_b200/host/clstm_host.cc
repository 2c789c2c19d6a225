// clstm_host.cc -- see clstm_host.h.  Host plumbing only; all arithmetic of the path runs behind the C ABI.
#include "clstm_host.h"
#include "clstm_extras.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>

namespace ocropus {

using std::string;
using std::vector;

void throwf(const char* fmt, ...) {  // utils.h:260-267: throws a const char* to a static buffer
  static thread_local char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  THROW((const char*)buf);
}
static void check(int rc) {
  if (rc) throwf("%s", clstm_b200_last_error());
}

// ------------------------------------------------------------------------------------------------ Assoc / Codec
Assoc::Assoc(const string& s) {  // "k=v:k=v"
  size_t start = 0;
  while (start <= s.size()) {
    size_t pos = s.find(':', start);
    string kv = s.substr(start, pos == string::npos ? string::npos : pos - start);
    size_t q = kv.find('=');
    if (q == string::npos) THROW("no '=' in Assoc");
    (*this)[kv.substr(0, q)] = kv.substr(q + 1);
    if (pos == string::npos) break;
    start = pos + 1;
    if (start >= s.size()) break;
  }
}
String Assoc::get(const string& key) const {
  auto it = find(key);
  if (it != end()) return it->second;
  if (super) return super->get(key);
  throwf("missing parameter: %s", key.c_str());
}
String Assoc::get(const string& key, String dflt) const {
  auto it = find(key);
  if (it != end()) return it->second;
  if (super) return super->get(key, dflt);
  return dflt;
}
void Codec::set(const vector<int>& data) {
  codec = data;
  encoder.clear();
  for (int i = 0; i < (int)codec.size(); i++) encoder.insert({codec[i], i});
}
wchar_t Codec::decode(int cls) const { return wchar_t(codec.at(cls)); }
std::wstring Codec::decode(const Classes& cs) const {
  std::wstring s;
  for (int c : cs) s.push_back(wchar_t(codec.at(c)));
  return s;
}
void Codec::encode(Classes& cs, const std::wstring& s) const {
  cs.clear();
  for (wchar_t ch : s) {
    auto it = encoder.find((int)ch);
    if (it == encoder.end()) throwf("character U+%04X is not in the codec", (unsigned)ch);
    if (it->second == 0) THROW("class 0 (blank) cannot be part of a transcript");
    cs.push_back(it->second);
  }
}

void Codec::build(const vector<string>& fnames, const std::wstring& extra) {   // clstm.cc:247-267
  std::set<int> codes;
  codes.insert(0);
  for (auto c : extra) codes.insert(int(c));
  for (const string& fname : fnames) {
    std::ifstream stream(fname);
    string line;
    while (getline(stream, line)) {
      if (line.empty() || line[0] == '#') continue;   // blank lines and comment lines carry no characters
      for (auto c : utf8_to_utf32(line)) codes.insert(int(c));
    }
  }
  set(vector<int>(codes.begin(), codes.end()));       // std::set iterates in increasing code order
}

// ------------------------------------------------------------------------------------------------ UTF-8
std::wstring utf8_to_utf32(const string& s) {
  std::wstring out;
  size_t i = 0;
  while (i < s.size()) {
    const unsigned c = (unsigned char)s[i];
    unsigned w;
    int extra;
    if (c < 0x80) { w = c; extra = 0; }
    else if ((c & 0xe0) == 0xc0) { w = c & 0x1f; extra = 1; }
    else if ((c & 0xf0) == 0xe0) { w = c & 0x0f; extra = 2; }
    else if ((c & 0xf8) == 0xf0) { w = c & 0x07; extra = 3; }
    else THROW("unicode character out of range");
    if (i + extra >= s.size() + (extra ? 0 : 1)) THROW("bad encoding");
    for (int k = 1; k <= extra; k++) w = (w << 6) | ((unsigned char)s[i + k] & 0x3f);
    out.push_back(wchar_t(w));
    i += 1 + extra;
  }
  return out;
}
string utf32_to_utf8(const std::wstring& s) {
  string out;
  for (wchar_t wc : s) {
    const unsigned c = (unsigned)wc;
    if (c < 0x80) out.push_back(char(c));
    else if (c <= 0x7ff) { out.push_back(char(0xc0 | (c >> 6))); out.push_back(char(0x80 | (c & 0x3f))); }
    else if (c <= 0xffff) {
      out.push_back(char(0xe0 | (c >> 12))); out.push_back(char(0x80 | ((c >> 6) & 0x3f))); out.push_back(char(0x80 | (c & 0x3f)));
    } else if (c <= 0x10ffff) {
      out.push_back(char(0xf0 | (c >> 18))); out.push_back(char(0x80 | ((c >> 12) & 0x3f)));
      out.push_back(char(0x80 | ((c >> 6) & 0x3f))); out.push_back(char(0x80 | (c & 0x3f)));
    } else THROW("unicode character out of range");
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ init RNG
namespace {
double& lcg_state() {  // batches.cc:11: seeded from env `seed`, default 0.1
  static double state = getenv("seed") ? atof(getenv("seed")) : 0.1;
  return state;
}
inline double randu() {  // batches.cc:13-17
  double& st = lcg_state();
  st = 189843.9384938 * st + 0.328340981343;
  st -= std::floor(st);
  return st;
}
}  // namespace
inline double randn_() {  // batches.cc:19-27: two LCG draws; the reference multiplies by r = -2 log(u1) itself (no square root)
  const double u1 = randu();
  const double u2 = randu();
  const double rr = -2 * std::log(u1);
  return rr * std::cos(2 * M_PI * u2);
}
void rinit(Params& m, int r, int c, Float s, const char* mode_, Float offset) {  // batches.cc:31-52 (draws: i outer, j inner)
  m.resize(r, c);
  const string mode(mode_);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      if (mode == "unif") m.v(i, j) = 2 * s * randu() - s + offset;
      else if (mode == "negbiased") m.v(i, j) = 3 * s * randu() - 2 * s + offset;
      else if (mode == "pos") m.v(i, j) = s * randu() + offset;
      else if (mode == "neg") m.v(i, j) = -s * randu() + offset;
      else if (mode == "normal") m.v(i, j) = s * randn_() + offset;
      else throwf("unsupported init_mode: %s", mode_);
    }
}
static void rinit_attr(Params& m, int r, int c, Assoc& attr) {  // clstm.cc:30-36
  const float s = (double)attr.get("init_scale", 0.01);
  const string mode = attr.get("init_mode", "negbiased");
  const float offset = (double)attr.get("init_offset", 0.0);
  rinit(m, r, c, s, mode.c_str(), offset);
}

// ------------------------------------------------------------------------------------------------ layers
namespace {
const char* kNotStandalone = "this layer only runs as part of the device-resident bidi network (clstm_b200)";

struct Container : INetwork {   // Parallel / Reversed: pure structure, evaluated inside the fused device net
  void forward() override { THROW(kNotStandalone); }
  void backward() override { THROW(kNotStandalone); }
  int ninput() override { return sub.empty() ? INetwork::ninput() : sub[0]->ninput(); }
};
struct Reversed : Container {
  int noutput() override { return sub[0]->noutput(); }          // clstm.cc:459
};
struct Parallel : Container {
  int noutput() override { return sub[0]->noutput() + sub[1]->noutput(); }   // clstm.cc:507-511
};
struct NPLSTM : INetwork {      // parameter container of GenericNPLSTM<SIG,TANH,TANH>  clstm.cc:546-599
  Params WGI, WGF, WGO, WCI;
  NPLSTM() { enroll(WGI, "WGI"); enroll(WGF, "WGF"); enroll(WGO, "WGO"); enroll(WCI, "WCI"); }
  void initialize() override {
    const int ni = (int)(double)attr.get("ninput"), no = (int)(double)attr.get("noutput");
    rinit_attr(WGI, no, ni + no + 1, attr);      // RNG draw order clstm.cc:588-591
    rinit_attr(WGF, no, ni + no + 1, attr);
    rinit_attr(WGO, no, ni + no + 1, attr);
    rinit_attr(WCI, no, ni + no + 1, attr);
  }
  int noutput() override { return WGI.rows(); }
  int ninput() override { return WGI.cols() - 1 - WGI.rows(); }   // postLoad clstm.cc:594-599
  void forward() override { THROW(kNotStandalone); }
  void backward() override { THROW(kNotStandalone); }
};
struct SoftmaxLayer : INetwork {  // clstm.cc:391-419
  Params W1;
  SoftmaxLayer() { enroll(W1, "W1"); }
  void initialize() override {
    const int no = (int)(double)attr.get("noutput"), ni = (int)(double)attr.get("ninput");
    if (no < 2) THROW("Softmax requires no>=2");
    rinit_attr(W1, no, ni + 1, attr);
  }
  int noutput() override { return W1.rows(); }
  int ninput() override { return W1.cols() - 1; }
  void forward() override { THROW(kNotStandalone); }
  void backward() override { THROW(kNotStandalone); }
};

struct FullLayer : INetwork {     // parameter container of Full<NONLIN> clstm.cc:354-389 (Linear/Sigmoid/Tanh/ReluLayer)
  Params W1;
  FullLayer() { enroll(W1, "W1"); }
  void initialize() override {
    const int no = (int)(double)attr.get("noutput"), ni = (int)(double)attr.get("ninput");
    rinit_attr(W1, no, ni + 1, attr);
  }
  int noutput() override { return W1.rows(); }
  int ninput() override { return W1.cols() - 1; }
  void forward() override { THROW(kNotStandalone); }
  void backward() override { THROW(kNotStandalone); }
};
int cell_code(const string& kind) {      // LSTM variants clstm.cc:655-668 -> clstm_b200_cfg_ex.cell, -1: not an LSTM
  static const char* k[5] = {"NPLSTM", "LINNPLSTM", "RELUTANHNPLSTM", "RELUNPLSTM", "RELU2NPLSTM"};
  for (int i = 0; i < 5; i++) if (kind == k[i]) return i;
  return -1;
}
int output_code(const string& kind) {    // output layers -> clstm_b200_cfg_ex.output, -2: not an output layer
  static const char* k[5] = {"SoftmaxLayer", "SigmoidLayer", "LinearLayer", "TanhLayer", "ReluLayer"};
  for (int i = 0; i < 5; i++) if (kind == k[i]) return i;
  return -2;
}

// Stacked{ recurrent block [, recurrent block] [, output layer] }: the whole tree runs as one device net.  A recurrent
// block is an LSTM, Reversed{LSTM} or Parallel{LSTM, Reversed{LSTM}} (lstm1 / revlstm1 / bidi / bidi2 / perplstm of
// clstm_prefab.cc:22-129; a Stacked with a single Stacked child is looked through).
struct Stacked : INetwork {      // clstm.cc:421-456
  clstm_b200_net* h = nullptr;
  bool weights_on_device = false;   // device copy current w.r.t. the host Params.v
  bool host_stale = false;          // device weights changed (sgd_update) since the last download
  int lastB = 0, lastT = 0;
  ~Stacked() override { clstm_b200_destroy(h); }
  int noutput() override { return sub.back()->noutput(); }
  int ninput() override { return sub[0]->ninput(); }
  // recognise one recurrent block; returns false if `net` is not one
  static bool parse_block(INetwork* net, int& direction, int& nhidden, int& cell) {
    auto lstm = [&](INetwork* l) {
      const int c = cell_code(l->kind);
      if (c < 0 || (cell >= 0 && c != cell)) return false;   // one cell type per network on the device
      cell = c;
      nhidden = l->noutput();
      return true;
    };
    if (cell_code(net->kind) >= 0) { direction = 0; return lstm(net); }
    if (net->kind == "Reversed" && net->sub.size() == 1) { direction = 1; return lstm(net->sub[0].get()); }
    if (net->kind == "Parallel" && net->sub.size() == 2 && net->sub[1]->kind == "Reversed" && net->sub[1]->sub.size() == 1) {
      direction = 2;
      int nh2 = 0;
      if (!lstm(net->sub[0].get())) return false;
      const int nh1 = nhidden;
      if (!lstm(net->sub[1]->sub[0].get())) return false;
      nh2 = nhidden;
      return nh1 == nh2;
    }
    return false;
  }
  bool describe(clstm_b200_cfg_ex& cfg) {
    INetwork* root = this;
    while (root->sub.size() == 1 && root->sub[0]->kind == "Stacked") root = root->sub[0].get();
    memset(&cfg, 0, sizeof cfg);
    cfg.ninput = ninput(); cfg.noutput = noutput();
    cfg.device = (int)(double)attr.get("gpu", 0);
    int cell = -1;
    size_t nsub = root->sub.size();
    cfg.output = -1;
    if (nsub >= 1 && output_code(root->sub[nsub - 1]->kind) >= 0) { cfg.output = output_code(root->sub[nsub - 1]->kind); nsub--; }
    if (nsub < 1 || nsub > 2) return false;
    cfg.nblocks = (int)nsub;
    for (size_t k = 0; k < nsub; k++)
      if (!parse_block(root->sub[k].get(), cfg.direction[k], cfg.nhidden[k], cell)) return false;
    cfg.cell = cell;
    return true;
  }
  void ensure_device() {
    if (h) return;
    clstm_b200_cfg_ex cfg;
    if (!describe(cfg))
      THROW("this topology does not run on the device (supported: lstm1, revlstm1, bidi, bidi2, perplstm; clstm_prefab.cc:22-129)");
    check(clstm_b200_create_ex(&cfg, &h));
    weights_on_device = false;
  }
  void upload() {
    ensure_device();
    if (weights_on_device) return;
    Network self(this, [](INetwork*) {});
    vector<Float> flat(n_params(self));
    // flatten the host tree directly (get_params would try to sync from the device first)
    size_t k = 0;
    walk_params(self, [&](const string&, Params* p) { for (Float x : p->v.data) flat[k++] = x; });
    check(clstm_b200_set_params(h, flat.data(), flat.size()));
    weights_on_device = true;
    host_stale = false;
  }
  void download() {               // device weights -> host Params.v (before saving / get_params)
    if (!h || !host_stale) return;
    Network self(this, [](INetwork*) {});
    vector<Float> flat(clstm_b200_nparams(h));
    check(clstm_b200_get_params(h, flat.data(), flat.size()));
    size_t k = 0;
    walk_params(self, [&](const string&, Params* p) { for (Float& x : p->v.data) x = flat[k++]; });
    host_stale = false;
  }
  void forward() override {        // Stacked::forward clstm.cc:424-439; B = inputs.cols() lines of equal length
    if (inputs.size() <= 0 || inputs.rows() <= 0 || inputs.cols() <= 0) THROW("empty inputs");
    upload();
    const int T = inputs.size(), ni = inputs.rows(), B = inputs.cols(), nc = noutput();
    if (ni != ninput()) throwf("input dimension %d does not match the network (%d)", ni, ninput());
    vector<float> x((size_t)B * T * ni), out((size_t)B * T * nc);
    for (int b = 0; b < B; b++)
      for (int t = 0; t < T; t++)
        for (int i = 0; i < ni; i++) x[((size_t)b * T + t) * ni + i] = inputs[t].v(i, b);
    vector<int> Ts(B, T);
    check(clstm_b200_forward(h, x.data(), Ts.data(), B, out.data()));
    outputs.resize(T, nc, B);
    for (int b = 0; b < B; b++)
      for (int t = 0; t < T; t++)
        for (int c = 0; c < nc; c++) outputs[t].v(c, b) = out[((size_t)b * T + t) * nc + c];
    lastB = B; lastT = T;
  }
  // CLSTMOCR::fwdbwd/predict prologue (clstmhl.h:201-205): normalizer->measure + normalize + set_inputs + forward with
  // the normalised line produced on the device; inputs/outputs are mirrored to the host Sequences afterwards.
  void forward_raw(Tensor2& raw, INormalizer& nm) {
    upload();
    const int ni = ninput(), nc = noutput();
    if (nm.target_height != ni) throwf("normalizer target_height %d does not match the network (%d)", nm.target_height, ni);
    int W = raw.dimension(0), H = raw.dimension(1), T = 0;
    float p[4];
    nm.abi_params(p);
    check(clstm_b200_normalize_batch(h, raw.ptr(), &W, &H, 1, nm.kind(), p, nullptr, nullptr, &T));
    vector<float> x((size_t)T * ni), out((size_t)T * nc);
    check(clstm_b200_forward_resident(h, out.data()));
    check(clstm_b200_get_inputs(h, x.data()));
    inputs.resize(T, ni, 1);
    outputs.resize(T, nc, 1);
    for (int t = 0; t < T; t++) {
      for (int i = 0; i < ni; i++) inputs[t].v(i, 0) = x[(size_t)t * ni + i];
      for (int c = 0; c < nc; c++) outputs[t].v(c, 0) = out[(size_t)t * nc + c];
    }
    lastB = 1; lastT = T;
  }
  void backward() override {       // Stacked::backward clstm.cc:440-454; outputs[t].d set by the caller
    if (!h || outputs.size() != lastT || outputs.cols() != lastB) THROW("backward called without a matching forward");
    const int T = lastT, B = lastB, nc = noutput(), ni = ninput();
    vector<float> dl((size_t)B * T * nc), din((size_t)B * T * ni);
    for (int b = 0; b < B; b++)
      for (int t = 0; t < T; t++)
        for (int c = 0; c < nc; c++) dl[((size_t)b * T + t) * nc + c] = outputs[t].d(c, b);
    check(clstm_b200_backward(h, dl.data(), din.data()));
    for (int b = 0; b < B; b++)
      for (int t = 0; t < T; t++)
        for (int i = 0; i < ni; i++) inputs[t].d(i, b) = din[((size_t)b * T + t) * ni + i];
  }
};

Stacked* as_device_root(INetwork* net) {
  auto* s = dynamic_cast<Stacked*>(net);
  return s;
}
Stacked* g_last_forward = nullptr;   // the net whose outputs the free-standing CTC functions refer to

template <class T>
void register_layer(const char* name) {  // clstm.cc:117-128
  string s(name);
  layer_factories[s] = [s]() {
    T* r = new T();
    r->kind = s;
    return r;
  };
}
struct Registrar {
  Registrar() {
    register_layer<Stacked>("Stacked");
    register_layer<Parallel>("Parallel");
    register_layer<Reversed>("Reversed");
    register_layer<NPLSTM>("NPLSTM");
    register_layer<NPLSTM>("LINNPLSTM");          // same parameter container, other cell nonlinearities on the device
    register_layer<NPLSTM>("RELUTANHNPLSTM");
    register_layer<NPLSTM>("RELUNPLSTM");
    register_layer<NPLSTM>("RELU2NPLSTM");
    register_layer<SoftmaxLayer>("SoftmaxLayer");
    register_layer<FullLayer>("LinearLayer");
    register_layer<FullLayer>("SigmoidLayer");
    register_layer<FullLayer>("TanhLayer");
    register_layer<FullLayer>("ReluLayer");
  }
};
}  // namespace

std::map<string, ILayerFactory> layer_factories;
static Registrar g_registrar;

Network make_layer(const string& kind) {
  Network net;
  auto it = layer_factories.find(kind);
  if (it != layer_factories.end()) net.reset(it->second());
  return net;
}
Network layer(const string& kind, int ninput, int noutput, const Assoc& args, const Networks& subs) {
  Network net = make_layer(kind);
  if (!net) {
    string accepted;
    for (auto& kv : layer_factories) accepted += kv.first + ",";
    throw std::string("unknown layer type:" + kind + ". Accepted layer kinds:" + accepted);   // clstm.cc:95-101 throws a std::string
  }
  for (auto& kv : args) net->attr.set(kv.first, kv.second);
  net->attr.set("ninput", ninput);
  net->attr.set("noutput", noutput);
  for (auto& s : subs) {
    net->add(s);
    s->attr.super = &net->attr;
  }
  net->initialize();
  return net;
}
// the 1-D prefabs of clstm_prefab.cc:22-129 (lstm_type / output_type as upstream; noutput == 1 defaults to a sigmoid)
namespace {
struct PrefabArgs {
  int ninput, nhidden, noutput;
  string lstm_type, output_type;
  explicit PrefabArgs(const Assoc& p) {
    ninput = (int)(double)p.get("ninput");
    noutput = (int)(double)p.get("noutput");
    nhidden = (int)(double)p.get("nhidden", noutput);
    lstm_type = p.get("lstm_type", "NPLSTM");
    output_type = p.get("output_type", noutput == 1 ? "SigmoidLayer" : "SoftmaxLayer");
  }
};
Network bidi_block(const Assoc& params, const string& lstm_type, int ninput, int nhidden) {
  return layer("Parallel", ninput, 2 * nhidden, {},
               {layer(lstm_type, ninput, nhidden, params, {}),
                layer("Reversed", ninput, ninput, {}, {layer(lstm_type, ninput, nhidden, params, {})})});
}
Network make_lstm1(const Assoc& params) {      // clstm_prefab.cc:22-32
  PrefabArgs a(params);
  return layer("Stacked", a.ninput, a.noutput, {},
               {layer(a.lstm_type, a.ninput, a.nhidden, params, {}), layer(a.output_type, a.nhidden, a.noutput, params, {})});
}
Network make_revlstm1(const Assoc& params) {   // clstm_prefab.cc:36-48
  PrefabArgs a(params);
  return layer("Stacked", a.ninput, a.noutput, {},
               {layer("Reversed", a.ninput, a.nhidden, {}, {layer(a.lstm_type, a.ninput, a.nhidden, params, {})}),
                layer(a.output_type, a.nhidden, a.noutput, params, {})});
}
Network make_bidi(const Assoc& params) {       // clstm_prefab.cc:52-68
  PrefabArgs a(params);
  return layer("Stacked", a.ninput, a.noutput, {},
               {bidi_block(params, a.lstm_type, a.ninput, a.nhidden), layer(a.output_type, 2 * a.nhidden, a.noutput, params, {})});
}
Network make_bidi0(const Assoc& params) {      // clstm_prefab.cc:72-82: no output layer, `noutput` hidden units per direction
  PrefabArgs a(params);
  return bidi_block(params, a.lstm_type, a.ninput, a.noutput);
}
Network make_bidi2(const Assoc& params) {      // clstm_prefab.cc:86-109
  PrefabArgs a(params);
  const int nhidden2 = (int)(double)params.get("nhidden2");
  return layer("Stacked", a.ninput, a.noutput, {},
               {bidi_block(params, a.lstm_type, a.ninput, a.nhidden), bidi_block(params, a.lstm_type, 2 * a.nhidden, nhidden2),
                layer(a.output_type, 2 * nhidden2, a.noutput, params, {})});
}
Network make_perplstm(const Assoc& params) {   // clstm_prefab.cc:111-125: a bidi net with a sigmoid output inside a Stacked
  PrefabArgs a(params);
  Assoc inner = {{"ninput", a.ninput}, {"nhidden", a.nhidden}, {"noutput", a.noutput},
                 {"output_type", String(params.get("output_type", "SigmoidLayer"))}};
  return layer("Stacked", a.ninput, a.noutput, {}, {make_bidi(inner)});
}
}  // namespace
Network make_net(const string& kind, const Assoc& args) {   // clstm_prefab.cc:163-173
  Network result;
  if (kind == "bidi") result = make_bidi(args);
  else if (kind == "lstm1") result = make_lstm1(args);
  else if (kind == "revlstm1") result = make_revlstm1(args);
  else if (kind == "bidi0") result = make_bidi0(args);
  else if (kind == "bidi2") result = make_bidi2(args);
  else if (kind == "perplstm") result = make_perplstm(args);
  else if (kind == "twod") THROW("the 2-D prefab `twod` is out of scope (SURVEY.md section 8)");
  else result = layer(kind, (int)(double)args.get("ninput"), (int)(double)args.get("noutput"), args, {});
  if (!result) throwf("no such network or layer: %s", kind.c_str());
  result->attr.set("kind", kind);
  return result;
}
Network make_net_init(const string& kind, const string& params) { return make_net(kind, Assoc(params)); }

void walk_params(Network net, ParamsFun f, const string& prefix) {
  for (auto& it : net->parameters) f(prefix + "." + it.first, it.second);
  for (auto& s : net->sub) walk_params(s, f, prefix + "." + s->kind);
}
int n_params(Network net) {
  int total = 0;
  walk_params(net, [&](const string&, Params* p) { total += p->v.total_size(); });
  return total;
}
void get_params(Network net, Float* params, int total) {
  if (auto* s = as_device_root(net.get())) s->download();
  int k = 0;
  walk_params(net, [&](const string&, Params* p) {
    if (k + p->v.total_size() > total) THROW("get_params size mismatch");
    for (Float x : p->v.data) params[k++] = x;
  });
  if (k != total) THROW("get_params size mismatch");
}
void set_params(Network net, const Float* params, int total) {
  int k = 0;
  walk_params(net, [&](const string&, Params* p) {
    if (k + p->v.total_size() > total) THROW("get_params size mismatch");   // sic: the reference reuses this message (clstm.cc:871)
    for (Float& x : p->v.data) x = params[k++];
  });
  if (k != total) THROW("get_params size mismatch");
  if (auto* s = as_device_root(net.get())) { s->weights_on_device = false; s->host_stale = false; }
}
void get_derivs(Network net, Float* params, int total) {
  auto* s = as_device_root(net.get());
  if (!s || !s->h) THROW("get_derivs: derivatives live on the device; run forward/backward first");
  if ((size_t)total != clstm_b200_nparams(s->h)) THROW("get_derivs size mismatch");
  check(clstm_b200_get_derivs(s->h, params, total));
}
void clear_derivs(Network net) {
  auto* s = as_device_root(net.get());
  if (s && s->h) check(clstm_b200_clear_derivs(s->h));
}
clstm_b200_net* device_handle(Network net) {
  auto* s = as_device_root(net.get());
  if (!s) THROW("not a device-resident network");
  s->upload();
  return s->h;
}

void set_inputs(Network net, Tensor2& image) {  // clstm.cc:684-690
  const int N = image.dimension(0), d = image.dimension(1);
  net->inputs.resize(N, d, 1);
  for (int t = 0; t < N; t++)
    for (int i = 0; i < d; i++) net->inputs[t].v(i, 0) = image(t, i);
}
void sgd_update(Network net) {  // clstm.cc:201-217: lr / momentum / clip from the attributes, update on the device
  auto* s = as_device_root(net.get());
  if (!s || !s->h) THROW("sgd_update: nothing to update (no forward/backward yet)");
  const Float lr = net->effective_lr();
  const Float momentum = (double)net->attr.get("momentum", 0.9);
  const Float gc = (double)net->attr.get("gradient_clip", 100.0);
  check(clstm_b200_sgd_update(s->h, lr, momentum, gc));
  s->host_stale = true;
}

// ------------------------------------------------------------------------------------------------ CTC glue
void mktargets(Sequence& seq, Classes& transcript, int ndim) {  // ctc.cc:148-157
  seq.resize(2 * (int)transcript.size() + 1, ndim, 1);
  for (int t = 0; t < seq.size(); t++) {
    if (t % 2 == 1) seq[t].v(transcript[(t - 1) / 2], 0) = 1;
    else seq[t].v(0, 0) = 1;
  }
}
static Stacked* ctc_net(Sequence& outputs) {
  Stacked* s = g_last_forward;
  if (!s || !s->h) THROW("ctc_align_targets: no device network has produced outputs yet");
  if (outputs.rows() != s->noutput()) THROW("ctc_align_targets: outputs do not belong to the device network");
  return s;
}
void ctc_align_targets(Sequence& posteriors, Sequence& outputs, Classes& targets) {  // ctc.cc:136-146
  if (outputs.cols() != 1) THROW("ctc_align_targets needs batch size 1");              // ctc.cc:116
  Stacked* s = ctc_net(outputs);
  const int T = outputs.size(), nc = outputs.rows(), S = (int)targets.size();
  vector<float> o((size_t)T * nc), al((size_t)T * nc);
  for (int t = 0; t < T; t++)
    for (int c = 0; c < nc; c++) o[(size_t)t * nc + c] = outputs[t].v(c, 0);
  check(clstm_b200_ctc_align_states(s->h, o.data(), &T, 1, targets.data(), &S, al.data()));
  posteriors.resize(T, nc, 1);
  for (int t = 0; t < T; t++)
    for (int c = 0; c < nc; c++) posteriors[t].v(c, 0) = al[(size_t)t * nc + c];
}
void ctc_align_targets(Sequence& posteriors, Sequence& outputs, Sequence& targets) {  // ctc.cc:114-134, one-hot states
  if (targets.cols() != 1 || outputs.rows() != targets.rows()) THROW("ctc_align_targets: shape mismatch");
  Classes states(targets.size());
  for (int s = 0; s < targets.size(); s++) {
    int hot = -1;
    for (int c = 0; c < targets.rows(); c++) {
      const Float v = targets[s].v(c, 0);
      if (v == 1) { if (hot >= 0) THROW("targets must be one-hot on the device path"); hot = c; }
      else if (v != 0) THROW("targets must be one-hot on the device path");
    }
    if (hot < 0) THROW("targets must be one-hot on the device path");
    states[s] = hot;
  }
  ctc_align_targets(posteriors, outputs, states);
}

// ------------------------------------------------------------------------------------------------ .clstm codec
// proto2 wire format of clstm.proto.  Writer emits repeated scalars unpacked like the C++
// protobuf writer does for proto2; the reader accepts both packed and unpacked.
namespace {
void put_varint(string& o, uint64_t v) {
  while (v >= 0x80) { o.push_back(char((v & 0x7f) | 0x80)); v >>= 7; }
  o.push_back(char(v));
}
void put_tag(string& o, int field, int wire) { put_varint(o, ((uint64_t)field << 3) | wire); }
void put_bytes(string& o, int field, const string& s) { put_tag(o, field, 2); put_varint(o, s.size()); o += s; }
void put_int32(string& o, int field, int v) { put_tag(o, field, 0); put_varint(o, (uint64_t)(int64_t)v); }
void put_float(string& o, int field, float f) { put_tag(o, field, 5); char b[4]; memcpy(b, &f, 4); o.append(b, 4); }

string encode_net(INetwork* net) {  // proto_of_net clstm_proto.cc:61-98
  if (net->kind == "") THROW("layer without kind cannot be saved");
  string o;
  put_bytes(o, 1, net->kind);
  put_int32(o, 10, net->ninput());
  put_int32(o, 11, net->noutput());
  for (int c : net->icodec.codec) put_int32(o, 12, c);
  for (int c : net->codec.codec) put_int32(o, 13, c);
  for (auto& kv : net->attr) {
    if (kv.first == "name" || kv.first == "ninput" || kv.first == "noutput") continue;
    string m;
    put_bytes(m, 1, kv.first);
    put_bytes(m, 2, kv.second);
    put_bytes(o, 20, m);
  }
  for (auto& it : net->parameters) {       // proto_of_params :35-45: dims then row-major values
    Params* a = it.second;
    string m;
    put_bytes(m, 1, it.first);
    put_int32(m, 2, a->rows());
    put_int32(m, 2, a->cols());
    for (int i = 0; i < a->rows(); i++)
      for (int j = 0; j < a->cols(); j++) put_float(m, 3, a->v(i, j));
    put_bytes(o, 30, m);
  }
  for (auto& s : net->sub) put_bytes(o, 40, encode_net(s.get()));
  return o;
}

struct Reader {
  const unsigned char* p;
  const unsigned char* e;
  bool ok = true;
  uint64_t varint() {
    uint64_t v = 0;
    for (int sh = 0; sh < 70; sh += 7) {
      if (p >= e) { ok = false; return 0; }
      const unsigned char b = *p++;
      v |= (uint64_t)(b & 0x7f) << sh;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  Reader sub() {
    const uint64_t n = varint();
    if (!ok || n > (uint64_t)(e - p)) { ok = false; return Reader{p, p}; }
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wire) {
    if (wire == 0) varint();
    else if (wire == 1) { if (e - p < 8) ok = false; else p += 8; }
    else if (wire == 2) sub();
    else if (wire == 5) { if (e - p < 4) ok = false; else p += 4; }
    else ok = false;
  }
};
string str_of(Reader r) { return string((const char*)r.p, (const char*)r.e); }

void decode_array(Reader r, string& name, vector<int>& dim, vector<float>& val) {
  while (r.ok && r.p < r.e) {
    const uint64_t tag = r.varint();
    const int field = (int)(tag >> 3), wire = (int)(tag & 7);
    if (field == 1 && wire == 2) name = str_of(r.sub());
    else if (field == 2 && wire == 0) dim.push_back((int)r.varint());
    else if (field == 2 && wire == 2) { Reader s = r.sub(); while (s.ok && s.p < s.e) dim.push_back((int)s.varint()); }
    else if (field == 3 && wire == 5) { float f; if (r.e - r.p < 4) { r.ok = false; break; } memcpy(&f, r.p, 4); r.p += 4; val.push_back(f); }
    else if (field == 3 && wire == 2) {
      Reader s = r.sub();
      while (s.p + 4 <= s.e) { float f; memcpy(&f, s.p, 4); s.p += 4; val.push_back(f); }
    } else r.skip(wire);
  }
  if (!r.ok) THROW("bad format (Array)");
}

constexpr int kMaxNetDepth = 16;   // the deepest reference topology nests 4 levels; a corrupt file must not overflow the stack
Network decode_net(Reader r, int depth = 0) {  // net_of_proto clstm_proto.cc:100-137
  if (depth > kMaxNetDepth) return Network();
  string kind;
  int ninput = -1, noutput = -1;
  vector<int> icodec, codec;
  vector<std::pair<string, string>> attrs;
  struct Arr { string name; vector<int> dim; vector<float> val; };
  vector<Arr> arrays;
  vector<Reader> subs;
  while (r.ok && r.p < r.e) {
    const uint64_t tag = r.varint();
    const int field = (int)(tag >> 3), wire = (int)(tag & 7);
    if (field == 1 && wire == 2) kind = str_of(r.sub());
    else if (field == 10 && wire == 0) ninput = (int)r.varint();
    else if (field == 11 && wire == 0) noutput = (int)r.varint();
    else if ((field == 12 || field == 13) && wire == 0) (field == 12 ? icodec : codec).push_back((int)r.varint());
    else if ((field == 12 || field == 13) && wire == 2) {
      Reader s = r.sub();
      while (s.ok && s.p < s.e) (field == 12 ? icodec : codec).push_back((int)s.varint());
    } else if (field == 20 && wire == 2) {
      Reader s = r.sub();
      string k, v;
      while (s.ok && s.p < s.e) {
        const uint64_t t2 = s.varint();
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) k = str_of(s.sub());
        else if ((t2 >> 3) == 2 && (t2 & 7) == 2) v = str_of(s.sub());
        else s.skip((int)(t2 & 7));
      }
      attrs.push_back({k, v});
    } else if (field == 30 && wire == 2) {
      Arr a;
      decode_array(r.sub(), a.name, a.dim, a.val);
      arrays.push_back(std::move(a));
    } else if (field == 40 && wire == 2) subs.push_back(r.sub());
    else r.skip(wire);
  }
  if (!r.ok || kind == "" || ninput < 0 || noutput < 0) return Network();
  Network net = make_layer(kind);
  if (!net) throwf("unknown layer kind in file: %s", kind.c_str());
  net->attr.set("ninput", ninput);
  net->attr.set("noutput", noutput);
  for (auto& kv : attrs) net->attr.set(kv.first, kv.second);
  net->icodec.set(icodec);
  net->codec.set(codec);
  for (auto& a : arrays) {                 // params_of_proto :47-59 (name lookup; derivatives start at zero)
    auto it = net->parameters.find(a.name);
    if (it == net->parameters.end()) throwf("unknown parameter in file: %s", a.name.c_str());
    if (a.dim.size() != 2) throwf("bad format (Mat, %s, %d)", a.name.c_str(), (int)a.dim.size());
    if (a.dim[0] < 0 || a.dim[1] < 0 || (long long)a.dim[0] * a.dim[1] > (1LL << 30)) THROW("bad size (Mat)");
    Params* p = it->second;
    p->resize(a.dim[0], a.dim[1]);
    if (!a.val.empty()) {
      if ((int)a.val.size() != a.dim[0] * a.dim[1]) THROW("bad size (Mat)");
      size_t k = 0;
      for (int i = 0; i < a.dim[0]; i++)
        for (int j = 0; j < a.dim[1]; j++) p->v(i, j) = a.val[k++];
    }
  }
  for (size_t i = 0; i < subs.size(); i++) {
    Network s = decode_net(subs[i], depth + 1);
    if (!s) return Network();
    net->add(s);
    net->sub[i]->attr.super = &net->attr;
  }
  net->postLoad();
  return net;
}
}  // namespace

bool write_as_proto(std::ostream& output, INetwork* net) {
  if (auto* s = as_device_root(net)) s->download();
  const string bytes = encode_net(net);
  output.write(bytes.data(), bytes.size());
  return (bool)output;
}
Network read_as_proto(std::istream& input) {
  std::stringstream ss;
  ss << input.rdbuf();
  const string bytes = ss.str();
  Reader r{(const unsigned char*)bytes.data(), (const unsigned char*)bytes.data() + bytes.size()};
  return decode_net(r);
}
bool save_as_proto(const string& fname, INetwork* net) {
  std::ofstream stream(fname, std::ios::binary);
  if (!stream) return false;
  return write_as_proto(stream, net);
}
Network load_as_proto(const string& fname) {
  std::ifstream stream(fname, std::ios::binary);
  if (!stream) throwf("cannot open: %s", fname.c_str());
  return read_as_proto(stream);
}
bool maybe_save_net(const string& file, Network net) { return save_as_proto(file, net.get()); }
Network maybe_load_net(const string& file) { return load_as_proto(file); }
void save_net(const string& file, Network net) {
  if (!save_as_proto(file, net.get())) THROW("error saving network");
}
Network load_net(const string& file) {
  Network result = load_as_proto(file);
  if (!result) THROW("error loading network");
  return result;
}

// ------------------------------------------------------------------------------------------------ CLSTMOCR
namespace {
void run_forward(CLSTMOCR& ocr, Tensor2& raw) {   // measure + normalize + set_inputs + forward (clstmhl.h:202-206)
  Stacked* s = as_device_root(ocr.net.get());
  if (!s) THROW("CLSTMOCR: the network root is not a device network");
  if (!ocr.normalizer) THROW("CLSTMOCR: no normalizer (createBidi or load first)");
  ocr.normalizer->target_height = ocr.target_height;
  s->forward_raw(raw, *ocr.normalizer);
  ocr.image.resize(s->lastT, s->ninput());
  for (int t = 0; t < s->lastT; t++)
    for (int i = 0; i < s->ninput(); i++) ocr.image(t, i) = s->inputs[t].v(i, 0);
  g_last_forward = s;
}
void device_decode(Network& net, int which, Classes& cs, vector<int>* where) {  // trivial_decode ctc.cc:159-194 on the device
  Stacked* s = as_device_root(net.get());
  const int cap = std::max(1, s->lastT / 2 + 1);
  vector<int> cls(cap), locs(cap);
  int count = 0;
  check(clstm_b200_decode(s->h, which, cls.data(), locs.data(), &count, cap));
  cs.assign(cls.begin(), cls.begin() + count);
  if (where) where->assign(locs.begin(), locs.begin() + count);
}
}  // namespace

bool CLSTMOCR::maybe_load(const string& fname) {
  net = maybe_load_net(fname);
  if (!net) {
    std::cerr << "WARNING: could not load CLSTMOCR net from " << fname;
    return false;
  }
  nclasses = net->codec.size();
  target_height = net->ninput();
  normalizer.reset(make_CenterNormalizer());
  normalizer->target_height = target_height;
  return true;
}
void CLSTMOCR::load(const string& fname) {
  if (!maybe_load(fname)) throw std::string("Could not load CLSTMOCR net from file: " + fname);
}
void CLSTMOCR::save(const string& fname) {
  if (!maybe_save(fname)) throw std::string("Could not save CLSTMOCR net to file: " + fname);
}
void CLSTMOCR::createBidi(const vector<int> codec, int nhidden) {
  nclasses = (int)codec.size();
  net = make_net("bidi", {{"ninput", target_height}, {"noutput", nclasses}, {"nhidden", nhidden}});
  net->initialize();   // no-op on the root, as upstream (clstmhl.h:196)
  net->codec.set(codec);
  normalizer.reset(make_CenterNormalizer());
  normalizer->target_height = target_height;
}
std::wstring CLSTMOCR::fwdbwd(Tensor2& raw, const std::wstring& target) {  // clstmhl.h:201-217, step by step
  run_forward(*this, raw);
  Classes transcript;
  net->codec.encode(transcript, target);
  mktargets(targets, transcript, nclasses);
  ctc_align_targets(aligned, net->outputs, targets);
  for (int t = 0; t < aligned.size(); t++)
    for (int c = 0; c < nclasses; c++) net->outputs[t].d(c, 0) = aligned[t].v(c, 0) - net->outputs[t].v(c, 0);
  net->backward();
  Classes outputs;
  device_decode(net, 0, outputs, nullptr);
  return net->codec.decode(outputs);
}
std::wstring CLSTMOCR::train(Tensor2& raw, const std::wstring& target) {
  std::wstring result = fwdbwd(raw, target);
  update();
  return result;
}
std::string CLSTMOCR::aligned_utf8() {   // clstmhl.h:224-229: decode of the aligned posteriors
  Stacked* s = as_device_root(net.get());
  if (!s || !s->h || aligned.size() == 0) THROW("aligned_utf8: no alignment yet");
  // the free-standing aligner left the posteriors (and their per-column argmax) on the device
  const int cap = std::max(1, aligned.size() / 2 + 1);
  vector<int> cls(cap), locs(cap);
  int count = 0;
  check(clstm_b200_fetch_decoded(s->h, 1, cls.data(), locs.data(), &count, cap));
  Classes cs(cls.begin(), cls.begin() + count);
  return utf32_to_utf8(net->codec.decode(cs));
}
std::wstring CLSTMOCR::predict(Tensor2& raw, vector<int>* where) {  // clstmhl.h:233-242
  run_forward(*this, raw);
  Classes outputs;
  device_decode(net, 0, outputs, where);
  return net->codec.decode(outputs);
}
void CLSTMOCR::predict(vector<CharPrediction>& preds, Tensor2& raw) {  // clstmhl.h:243-261
  run_forward(*this, raw);
  Classes outputs;
  vector<int> where;
  device_decode(net, 0, outputs, &where);
  preds.clear();
  for (int i = 0; i < (int)outputs.size(); i++) {
    const int t = where[i], cls = outputs[i];
    preds.push_back(CharPrediction{i, t, net->codec.decode(cls), net->outputs[t].v(cls, 0)});
  }
}
void CLSTMOCR::get_outputs(Tensor2& outputs) {  // clstmhl.h:265-271
  Sequence& o = net->outputs;
  outputs.resize(o.size(), o.rows());
  for (int t = 0; t < outputs.dimension(0); t++)
    for (int c = 0; c < outputs.dimension(1); c++) outputs(t, c) = o[t].v(c, 0);
}
namespace {
struct RawBatch {   // B raw lines + transcripts marshalled for clstm_b200_normalize_batch / prefetch_raw_batch
  vector<int> W, H, T, L, labels;
  vector<float> raw;
  float p[4];
  int kind;
};
void marshal(CLSTMOCR& ocr, std::vector<Tensor2>& images, const std::vector<std::wstring>& tg, RawBatch& rb) {
  if (images.size() != tg.size() || images.empty()) THROW("train_batch: need one transcript per image");
  if (!ocr.normalizer) THROW("CLSTMOCR: no normalizer (createBidi or load first)");
  const int B = (int)images.size();
  rb.W.resize(B); rb.H.resize(B); rb.T.assign(B, 0); rb.L.resize(B); rb.labels.clear(); rb.raw.clear();
  for (int b = 0; b < B; b++) {
    rb.W[b] = images[b].dimension(0); rb.H[b] = images[b].dimension(1);
    Classes cs;
    ocr.net->codec.encode(cs, tg[b]);
    rb.L[b] = (int)cs.size();
    rb.labels.insert(rb.labels.end(), cs.begin(), cs.end());
    rb.raw.insert(rb.raw.end(), images[b].data.begin(), images[b].data.end());
  }
  if (rb.labels.empty()) rb.labels.push_back(0);   // never dereferenced with all-empty transcripts, but must not be null
  ocr.normalizer->target_height = ocr.target_height;
  ocr.normalizer->abi_params(rb.p);
  rb.kind = ocr.normalizer->kind();
}
std::vector<std::wstring> decoded(CLSTMOCR& ocr, clstm_b200_net* h, int B, int cap) {
  vector<int> cls((size_t)B * cap), locs((size_t)B * cap), cnt(B);
  check(clstm_b200_fetch_decoded(h, 0, cls.data(), locs.data(), cnt.data(), cap));
  std::vector<std::wstring> out(B);
  for (int b = 0; b < B; b++) {
    Classes cs(cls.begin() + (size_t)b * cap, cls.begin() + (size_t)b * cap + cnt[b]);
    out[b] = ocr.net->codec.decode(cs);
  }
  return out;
}
Stacked* device_root_of(CLSTMOCR& ocr) {
  Stacked* s = as_device_root(ocr.net.get());
  if (!s) THROW("CLSTMOCR: the network root is not a device network");
  s->upload();
  return s;
}
}  // namespace

std::vector<std::wstring> CLSTMOCR::train_batch(std::vector<Tensor2>& images, const std::vector<std::wstring>& tg) {
  RawBatch rb;
  marshal(*this, images, tg, rb);
  Stacked* s = device_root_of(*this);
  const int B = (int)images.size();
  check(clstm_b200_normalize_batch(s->h, rb.raw.data(), rb.W.data(), rb.H.data(), B, rb.kind, rb.p, rb.labels.data(),
                                   rb.L.data(), rb.T.data()));
  const Float lr = net->effective_lr();
  const Float momentum = (double)net->attr.get("momentum", 0.9), gc = (double)net->attr.get("gradient_clip", 100.0);
  check(clstm_b200_step_resident(s->h, lr, momentum, gc));
  s->host_stale = true;
  int tmax = 0;
  for (int b = 0; b < B; b++) tmax = std::max(tmax, rb.T[b]);
  return decoded(*this, s->h, B, tmax / 2 + 1);
}
void CLSTMOCR::prefetch_batch(std::vector<Tensor2>& images, const std::vector<std::wstring>& tg) {
  RawBatch rb;
  marshal(*this, images, tg, rb);
  Stacked* s = device_root_of(*this);
  const int B = (int)images.size();
  // returns after the normaliser's round trip on the copy stream; the raw pixels have been consumed by then
  check(clstm_b200_prefetch_raw_batch(s->h, rb.raw.data(), rb.W.data(), rb.H.data(), B, rb.kind, rb.p, rb.labels.data(),
                                      rb.L.data(), rb.T.data()));
  int tmax = 0;
  for (int b = 0; b < B; b++) tmax = std::max(tmax, rb.T[b]);
  next_B = B; next_cap = tmax / 2 + 1;
}
void CLSTMOCR::train_prefetched() {
  Stacked* s = device_root_of(*this);
  if (next_B <= 0) THROW("train_prefetched: no prefetched batch");
  const Float lr = net->effective_lr();
  const Float momentum = (double)net->attr.get("momentum", 0.9), gc = (double)net->attr.get("gradient_clip", 100.0);
  check(clstm_b200_step_prefetched(s->h, lr, momentum, gc));
  s->host_stale = true;
  pipe_B = next_B; pipe_cap = next_cap;
  next_B = 0;
}
std::vector<std::wstring> CLSTMOCR::fetch_results() {
  Stacked* s = device_root_of(*this);
  if (pipe_B <= 0) THROW("fetch_results: no step in flight");
  return decoded(*this, s->h, pipe_B, pipe_cap);
}

}  // namespace ocropus
