// clstm_host.h -- host-side C++ mirror of the reference interface for the hot path.
//
// Same names, argument meaning and error behaviour as tmbdev/clstm (paths relative to /root/reference):
//   Float/Tensor2/Batch/Params/Sequence        tensor.h:62-66,176-335 ; batches.h:12-148
//   Assoc, Codec, INetwork, Network            clstm.h:35-152
//   layer_factories / make_layer / layer / make_net / make_net_init      clstm.cc:79-128, clstm_prefab.cc:163-189
//   set_inputs, sgd_update, n_params, get/set_params, get_derivs, clear_derivs    clstm.cc:684-690, 201-217, 838-918
//   mktargets, ctc_align_targets, trivial_decode                          clstm.h:310-320, ctc.cc
//   save_net/load_net/maybe_*  (.clstm proto2 files)                      clstm_proto.cc:35-181, clstm.proto
//   CLSTMOCR                                                              clstmhl.h:146-272
// Everything numerical happens on the GPU behind include/clstm_b200.h; this file is plumbing: host containers,
// the layer tree as parameter container (so files keep the reference's kind strings and parameter names), the
// hand-written proto2 wire codec (no protoc/libprotobuf on the box), and the CLSTMOCR wrapper.
// Errors: THROW(const char*) exactly like the reference (utils.h:260-267); C ABI failures are re-thrown with
// clstm_b200_last_error().
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

extern "C" {
#include "../../include/clstm_b200.h"
}

namespace ocropus {

typedef float Float;
typedef std::vector<int> Classes;

[[noreturn]] void throwf(const char* fmt, ...);
#define THROW(X) throw(X)

// ---- storage ---------------------------------------------------------------------------------------------------
struct Tensor2 {  // column-major, (i,j) at ptr[i + j*rows]  (tensor.h:252,288)
  int n = 0, m = 0;
  std::vector<Float> data;
  void resize(int r, int c) { n = r; m = c; data.assign((size_t)r * c, 0); }
  void setZero() { std::fill(data.begin(), data.end(), Float(0)); }
  int rows() const { return n; }
  int cols() const { return m; }
  int dimension(int i) const { return i == 0 ? n : m; }
  int total_size() const { return n * m; }
  Float& operator()(int i, int j) { return data[(size_t)i + (size_t)j * n]; }
  Float operator()(int i, int j) const { return data[(size_t)i + (size_t)j * n]; }
  Float* ptr() { return data.data(); }
};
struct Batch {  // batches.h:12-25
  Tensor2 v, d;
  int rows() const { return v.n; }
  int cols() const { return v.m; }
  void resize(int n, int m) { v.resize(n, m); d.resize(n, m); }
  void zeroGrad() { d.resize(v.n, v.m); }
};
typedef Batch Params;
struct Sequence {  // batches.h:45-148 (host only; resize zero-fills like batches.h:127)
  std::vector<Batch> steps;
  int nrows = 0, ncols = 0;
  int size() const { return (int)steps.size(); }
  int rows() const { return nrows; }
  int cols() const { return ncols; }
  void resize(int N, int n, int m) {
    nrows = n; ncols = m;
    steps.resize(N);
    for (auto& s : steps) s.resize(n, m);
  }
  void like(const Sequence& o) { resize(o.size(), o.rows(), o.cols()); }
  void clear() { steps.clear(); nrows = ncols = 0; }
  Batch& operator[](int i) { return steps[i]; }
  const Batch& operator[](int i) const { return steps[i]; }
  void zeroGrad() { for (auto& s : steps) s.zeroGrad(); }
};

// ---- attributes, codec -----------------------------------------------------------------------------------------
class String : public std::string {  // clstm.h:38-49
 public:
  String() {}
  String(const char* s) : std::string(s) {}
  String(const std::string& s) : std::string(s) {}
  String(int x) : std::string(std::to_string(x)) {}
  String(double x) : std::string(std::to_string(x)) {}
  operator double() const { return atof(c_str()); }
};
class Assoc : public std::map<std::string, String> {  // clstm.h:52-79
 public:
  Assoc() {}
  Assoc(const std::string& s);  // "k=v:k=v"  (clstm.cc:38-57)
  Assoc(std::initializer_list<std::pair<const std::string, String>> l) : std::map<std::string, String>(l) {}
  Assoc* super = nullptr;
  String get(const std::string& key) const;
  String get(const std::string& key, String dflt) const;
  void set(const std::string& key, String value) { (*this)[key] = value; }
};
class Codec {  // clstm.h:82-92, clstm.cc:219-267
 public:
  std::vector<int> codec;
  std::map<int, int> encoder;
  int size() const { return (int)codec.size(); }
  void set(const std::vector<int>& data);
  wchar_t decode(int cls) const;
  std::wstring decode(const Classes& cs) const;
  void encode(Classes& cs, const std::wstring& s) const;
  void build(const std::vector<std::string>& fnames, const std::wstring& extra);   // clstm.cc:247-267
};

// ---- networks --------------------------------------------------------------------------------------------------
class INetwork;
typedef std::shared_ptr<INetwork> Network;
typedef std::vector<Network> Networks;

class INetwork {  // clstm.h:98-152
 public:
  virtual ~INetwork() {}
  std::string kind = "";
  std::vector<Network> sub;
  std::map<std::string, Params*> parameters;
  virtual void add(Network net) { sub.push_back(net); }
  void enroll(Params& p, const char* name) { parameters[name] = &p; }
  Assoc attr;
  Sequence inputs, outputs;
  virtual int ninput() { return (int)(double)attr.get("ninput"); }
  virtual int noutput() { return (int)(double)attr.get("noutput"); }
  virtual void forward() = 0;
  virtual void backward() = 0;
  virtual void initialize() {}
  virtual void postLoad() {}
  virtual void setLearningRate(Float lr, Float momentum) {  // clstm.cc:163-166
    attr.set("learning_rate", (double)lr);
    attr.set("momentum", (double)momentum);
  }
  Float effective_lr() { return (Float)(double)attr.get("learning_rate"); }  // normalisation is a no-op upstream
  Codec codec, icodec;
};

typedef std::function<INetwork*(void)> ILayerFactory;
extern std::map<std::string, ILayerFactory> layer_factories;                 // clstm.cc:79
Network make_layer(const std::string& kind);                                 // clstm.cc:81-86 (empty on unknown kind)
Network layer(const std::string& kind, int ninput, int noutput, const Assoc& args, const Networks& subs);  // :88-115
Network make_net(const std::string& kind, const Assoc& params);              // clstm_prefab.cc:163-173 ("bidi")
Network make_net_init(const std::string& kind, const std::string& params);   // :178-188

typedef std::function<void(const std::string&, Params*)> ParamsFun;
void walk_params(Network net, ParamsFun f, const std::string& prefix = "");  // clstm.cc:59-62
int n_params(Network net);                                                   // clstm.cc:838-843
void get_params(Network net, Float* params, int total);                      // host copies of the tree's Params
void set_params(Network net, const Float* params, int total);
void get_derivs(Network net, Float* params, int total);
void clear_derivs(Network net);

// deterministic init (batches.cc:11-52), seed from env `seed` (default 0.1)
void rinit(Params& m, int r, int c, Float s, const char* mode, Float offset = 0);

// inputs: image is T x d with image(t,i) (Tensor2 col-major), batch 1   (clstm.cc:684-690)
void set_inputs(Network net, Tensor2& image);
void sgd_update(Network net);                                                // clstm.cc:201-217, on the device

// CTC (device).  outputs must be the `outputs` of the network that just ran forward (they live on the device).
void mktargets(Sequence& seq, Classes& transcript, int ndim);                // ctc.cc:148-157 (host one-hot container)
void ctc_align_targets(Sequence& posteriors, Sequence& outputs, Classes& targets);   // ctc.cc:136-146 (explicit state classes)
void ctc_align_targets(Sequence& posteriors, Sequence& outputs, Sequence& targets);  // ctc.cc:114-134 (one-hot states)

// persistence (.clstm = proto2 NetworkProto, clstm.proto)
bool write_as_proto(std::ostream& output, INetwork* net);
Network read_as_proto(std::istream& input);
bool save_as_proto(const std::string& fname, INetwork* net);
Network load_as_proto(const std::string& fname);
void save_net(const std::string& file, Network net);                         // throws "error saving network"
Network load_net(const std::string& file);                                   // throws "error loading network"
bool maybe_save_net(const std::string& file, Network net);
Network maybe_load_net(const std::string& file);

std::wstring utf8_to_utf32(const std::string& s);                            // pstring.h
std::string utf32_to_utf8(const std::wstring& s);

// The device-resident bidi net.  Registered as "Stacked" holder of the reference tree so that files carry the
// reference kind strings; exposes the handle for batched training.
clstm_b200_net* device_handle(Network net);

struct CharPrediction { int i; int x; wchar_t c; float p; };                 // clstmhl.h:18-23

struct INormalizer;                                                          // clstm_extras.h (extras.h:31-43)

// clstmhl.h:146-272.  `raw` images are Tensor2 raw(i, j), i = column, j = row, ink = 1 (clstmocrtrain.cc:74-75); they
// are measured and normalised ON THE DEVICE by `normalizer` (CenterNormalizer by default, like the reference) straight
// into the network's input batch.  With a NoNormalizer the image must already be target_height rows high.
struct CLSTMOCR {
  std::shared_ptr<INormalizer> normalizer;
  Network net;
  int target_height = 48;
  int nclasses = -1;
  Sequence aligned, targets;
  Tensor2 image;                                                             // the normalised line of the last call
  void setLearningRate(float lr, float mom) { net->setLearningRate(lr, mom); }
  bool maybe_load(const std::string& fname);
  void load(const std::string& fname);
  void save(const std::string& fname);
  bool maybe_save(const std::string& fname) { return maybe_save_net(fname, net); }
  void createBidi(const std::vector<int> codec, int nhidden);
  std::wstring fwdbwd(Tensor2& image, const std::wstring& target);
  void update() { sgd_update(net); }
  std::wstring train(Tensor2& image, const std::wstring& target);
  std::string train_utf8(Tensor2& image, const std::string& target) { return utf32_to_utf8(train(image, utf8_to_utf32(target))); }
  std::string aligned_utf8();
  std::wstring predict(Tensor2& image, std::vector<int>* where = 0);
  void predict(std::vector<CharPrediction>& preds, Tensor2& image);
  std::string predict_utf8(Tensor2& image) { return utf32_to_utf8(predict(image)); }
  void get_outputs(Tensor2& outputs);
  // minibatch extension: B raw lines -> normalise + forward + CTC + backward + update in one device step
  // (clstm_b200_normalize_batch + clstm_b200_step_resident); returns the decoded strings
  std::vector<std::wstring> train_batch(std::vector<Tensor2>& images, const std::vector<std::wstring>& targets);
  // the same as a two-deep input pipeline: prefetch_batch(i+1) normalises the next batch on the copy stream while the
  // step on batch i is still running; train_prefetched() launches the step on the prefetched batch and returns at once;
  // fetch_results() waits for it and returns its decoded strings.  Loop:  prefetch(0); { train_prefetched();
  // prefetch(next); fetch_results(); } ...
  void prefetch_batch(std::vector<Tensor2>& images, const std::vector<std::wstring>& targets);
  void train_prefetched();
  std::vector<std::wstring> fetch_results();
  int pipe_B = 0, pipe_cap = 0, next_B = 0, next_cap = 0;
};

}  // namespace ocropus
