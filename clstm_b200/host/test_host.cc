// test_host.cc -- exercises the host mirror the way the reference's own tests exercise its API
// (test-lstm.cc:75-153: build, train, save -> load round trip, parameter access; test-ocr.sh: train on one line until
// it is read back).  `test_host cpu <file>`: no device needed.  `test_host gpu`: needs a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "clstm_host.h"

using namespace ocropus;
using std::string;
using std::vector;

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) { fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); exit(1); } \
  } while (0)

static vector<int> demo_codec() {  // class 0 = blank, then 'a'..'e', ' '
  return {0, ' ', 'a', 'b', 'c', 'd', 'e'};
}
// a synthetic "line": each character lights a band of rows for 8 columns, 3 blank columns between characters
static Tensor2 render(const std::wstring& text, int H) {
  const int T = 6 + 11 * (int)text.size();
  Tensor2 img;
  img.resize(T, H);
  for (size_t k = 0; k < text.size(); k++) {
    const int band = (text[k] == ' ') ? 6 : (int)(text[k] - 'a');
    for (int t = 4 + 11 * (int)k; t < 12 + 11 * (int)k; t++)
      for (int i = 6 * band; i < 6 * band + 6 && i < H; i++) img(t, i) = 1.0f;
  }
  return img;
}

static int run_cpu(const char* fname) {
  // registry behaviour (clstm.cc:81-101)
  CHECK(!make_layer("NoSuchLayer"));
  bool threw = false;
  try { layer("NoSuchLayer", 1, 1, {}, {}); } catch (const std::string& s) { threw = s.find("Accepted layer kinds:") != string::npos; }
  CHECK(threw);
  threw = false;
  try { Assoc bad("novalue"); } catch (const char*) { threw = true; }
  CHECK(threw);
  Assoc a("x=1:y=two");
  CHECK((int)(double)a.get("x") == 1 && string(a.get("y")) == "two" && string(a.get("z", "dflt")) == "dflt");

  CLSTMOCR ocr;
  ocr.target_height = 48;
  ocr.createBidi(demo_codec(), 10);
  Network net = ocr.net;
  CHECK(net->kind == "Stacked" && net->sub.size() == 2 && net->sub[0]->kind == "Parallel");
  CHECK(net->ninput() == 48 && net->noutput() == 7);
  const int P = n_params(net);
  CHECK(P == 2 * 4 * 10 * (1 + 48 + 10) + 7 * (1 + 20));
  vector<Float> p0(P);
  get_params(net, p0.data(), P);
  for (Float v : p0) CHECK(v >= -0.02f - 1e-7f && v <= 0.01f + 1e-7f);   // negbiased, scale 0.01
  vector<string> names;
  walk_params(net, [&](const string& n, Params*) { names.push_back(n); });
  CHECK(names.size() == 9 && names[0] == ".Parallel.NPLSTM.WCI" && names[3] == ".Parallel.NPLSTM.WGO" &&
        names[4] == ".Parallel.Reversed.NPLSTM.WCI" && names[8] == ".SoftmaxLayer.W1");
  net->setLearningRate(1e-3, 0.9);
  net->attr.set("trial", 1234);
  ocr.save(fname);
  CLSTMOCR ocr2;
  ocr2.load(fname);
  CHECK(ocr2.nclasses == 7 && ocr2.target_height == 48);
  CHECK(n_params(ocr2.net) == P);
  vector<Float> p1(P);
  get_params(ocr2.net, p1.data(), P);
  CHECK(memcmp(p0.data(), p1.data(), P * sizeof(Float)) == 0);
  CHECK((int)(double)ocr2.net->attr.get("trial") == 1234);
  CHECK(std::fabs((double)ocr2.net->attr.get("learning_rate") - 1e-3) < 1e-9);
  CHECK(ocr2.net->codec.codec == demo_codec());
  CHECK(utf32_to_utf8(utf8_to_utf32("a\xc3\xa9\xe2\x82\xac")) == "a\xc3\xa9\xe2\x82\xac");
  // no device in this mode: the numerical entry points must fail loudly, not fall back
  threw = false;
  Tensor2 img = render(L"ab", 48);
  try { ocr.predict(img); } catch (const char* msg) { threw = string(msg).find("no CPU fallback") != string::npos || string(msg).find("device") != string::npos; }
  if (getenv("EXPECT_NO_GPU")) CHECK(threw);
  printf("host cpu ok: %d params saved to %s\n", P, fname);
  return 0;
}

static int run_gpu() {
  CLSTMOCR ocr;
  ocr.createBidi(demo_codec(), 16);
  ocr.setLearningRate(1e-2, 0.9);
  const std::wstring text = L"abc de";
  Tensor2 img = render(text, 48);
  std::wstring got;
  int it = 0;
  for (; it < 600; it++) {
    got = ocr.train(img, text);
    if (got == text && it > 20) break;
  }
  printf("trained %d steps, reads: %s\n", it, utf32_to_utf8(got).c_str());
  CHECK(ocr.predict(img) == text);
  CHECK(ocr.aligned_utf8() == utf32_to_utf8(text));
  vector<int> where;
  ocr.predict(img, &where);
  CHECK(where.size() == text.size());
  for (size_t k = 0; k < where.size(); k++) CHECK(where[k] >= 4 + 11 * (int)k - 2 && where[k] < 12 + 11 * (int)k + 4);
  vector<CharPrediction> preds;
  ocr.predict(preds, img);
  CHECK(preds.size() == text.size() && preds[0].c == L'a' && preds[0].p > 0.5f);
  Tensor2 outs;
  ocr.get_outputs(outs);
  CHECK(outs.dimension(0) == img.dimension(0) && outs.dimension(1) == 7);
  // save -> load: the reloaded net must read the same (test-lstm.cc:108-118)
  ocr.save("/tmp/clstm_b200_host_test.clstm");
  CLSTMOCR ocr2;
  ocr2.load("/tmp/clstm_b200_host_test.clstm");
  CHECK(ocr2.predict(img) == text);
  // derivatives and the momentum buffer are not saved (clstm_proto.cc:51): a step after reload works from zero
  ocr2.setLearningRate(1e-2, 0.9);
  ocr2.train(img, text);
  // zeroing the parameters breaks it, restoring fixes it (test-lstm.cc:129-150)
  const int P = n_params(ocr2.net);
  vector<Float> keep(P), zeros(P, 0.f);
  get_params(ocr2.net, keep.data(), P);
  set_params(ocr2.net, zeros.data(), P);
  CHECK(ocr2.predict(img) != text);
  set_params(ocr2.net, keep.data(), P);
  CHECK(ocr2.predict(img) == text);
  // minibatch step through the fused device call
  vector<Tensor2> imgs = {render(L"ab", 48), render(L"cde a", 48), render(text, 48)};
  vector<std::wstring> tgs = {L"ab", L"cde a", text};
  for (int k = 0; k < 300; k++) ocr.train_batch(imgs, tgs);
  auto res = ocr.train_batch(imgs, tgs);
  CHECK(res[0] == tgs[0] && res[1] == tgs[1] && res[2] == tgs[2]);
  printf("host gpu ok\n");
  return 0;
}

// loads a .clstm written by anybody (e.g. python-protobuf) and prints what it found, for cross-checking the reader
static int run_load(const char* fname) {
  Network net = load_net(fname);
  const int P = n_params(net);
  vector<Float> p(P);
  get_params(net, p.data(), P);
  double sum = 0;
  for (Float v : p) sum += v;
  printf("kind=%s ninput=%d noutput=%d nparams=%d sum=%.9g codec=%d lr=%s\n", net->kind.c_str(), net->ninput(),
         net->noutput(), P, sum, net->codec.size(), string(net->attr.get("learning_rate", "none")).c_str());
  for (int i = 0; i < 5 && i < P; i++) printf("p[%d]=%.9g\n", i, p[i]);
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && !strcmp(argv[1], "load")) return run_load(argv[2]);
    if (argc >= 3 && !strcmp(argv[1], "cpu")) return run_cpu(argv[2]);
    if (argc >= 2 && !strcmp(argv[1], "gpu")) return run_gpu();
    fprintf(stderr, "usage: test_host cpu <out.clstm> | test_host gpu\n");
    return 2;
  } catch (const char* msg) {
    fprintf(stderr, "FATAL: %s\n", msg);   // the reference CLIs catch const char* the same way (clstmocrtrain.cc:221-224)
    return 1;
  } catch (const std::string& msg) {
    fprintf(stderr, "FATAL: %s\n", msg.c_str());
    return 1;
  }
}
