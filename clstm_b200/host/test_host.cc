// test_host.cc -- exercises the host mirror the way the reference's own tests exercise its API
// (test-lstm.cc:75-153: build, train, save -> load round trip, parameter access; test-ocr.sh: train on one line until
// it is read back).  `test_host cpu <file>`: no device needed.  `test_host gpu`: needs a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "clstm_host.h"
#include "clstm_extras.h"

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
  // ---- the other 1-D prefabs and layer variants (clstm_prefab.cc:22-129, clstm.cc:382-389, 655-668)
  {
    Network b2 = make_net("bidi2", {{"ninput", 48}, {"nhidden", 5}, {"nhidden2", 7}, {"noutput", 9}});
    CHECK(b2->kind == "Stacked" && b2->sub.size() == 3 && b2->sub[1]->kind == "Parallel" && b2->sub[2]->kind == "SoftmaxLayer");
    CHECK(n_params(b2) == 2 * 4 * 5 * (1 + 48 + 5) + 2 * 4 * 7 * (1 + 10 + 7) + 9 * (1 + 14));
    Network l1 = make_net("lstm1", {{"ninput", 8}, {"nhidden", 4}, {"noutput", 1}});
    CHECK(l1->sub.size() == 2 && l1->sub[0]->kind == "NPLSTM" && l1->sub[1]->kind == "SigmoidLayer");   // noutput == 1 => sigmoid
    Network rl = make_net("revlstm1", {{"ninput", 8}, {"nhidden", 4}, {"noutput", 3}, {"lstm_type", "LINNPLSTM"}});
    CHECK(rl->sub[0]->kind == "Reversed" && rl->sub[0]->sub[0]->kind == "LINNPLSTM" && rl->sub[1]->kind == "SoftmaxLayer");
    Network b0 = make_net("bidi0", {{"ninput", 8}, {"noutput", 3}});
    CHECK(b0->kind == "Parallel" && b0->noutput() == 6 && n_params(b0) == 2 * 4 * 3 * (1 + 8 + 3));
    Network pp = make_net("perplstm", {{"ninput", 8}, {"nhidden", 4}, {"noutput", 2}});
    CHECK(pp->sub.size() == 1 && pp->sub[0]->kind == "Stacked" && pp->sub[0]->sub[1]->kind == "SigmoidLayer");
    threw = false;
    try { make_net("twod", {{"ninput", 8}, {"nhidden", 4}, {"noutput", 2}}); } catch (const char*) { threw = true; }
    CHECK(threw);
    const string f2 = string(fname) + ".bidi2";
    save_net(f2, b2);
    Network b2l = load_net(f2);
    CHECK(b2l->sub.size() == 3 && n_params(b2l) == n_params(b2));
    vector<Float> q0(n_params(b2)), q1(n_params(b2));
    get_params(b2, q0.data(), (int)q0.size());
    get_params(b2l, q1.data(), (int)q1.size());
    CHECK(q0 == q1);
  }
  // ---- utils.h / clstm.h helpers
  std::wstring k1 = L"kitten", k2 = L"sitting", e0 = L"";
  CHECK(levenshtein(k1, k2) == 3 && levenshtein(k2, k1) == 3 && levenshtein(k1, e0) == 6 && levenshtein(k1, k1) == 0);
  CHECK(ocropus::basename(string("/a/b.c/line-01.bin.png")) == "/a/b.c/line-01" && ocropus::basename(string("x")) == "x");
  {
    Trigger t(10, 35, 0);                    // every 10, forced at upto-1 = 34
    vector<int> fired;
    for (int i = 0; i < 35; i++) if (t(i)) fired.push_back(i);
    CHECK((fired == vector<int>{0, 10, 20, 30, 34}));
    Trigger u(10, -1, 0);
    u.skip0();
    CHECK(!u(0) && !u(9) && u(10) && u.since() == 10 && !u(11));
    Trigger z(0, 100, 0);
    CHECK(!z(5));
  }
  {
    const string gt = string(fname) + ".gt.txt";
    write_text(gt, string("# comment\nhello\n\nw\xc3\xb6rld"));
    Codec c;
    c.build({gt}, L"~");
    CHECK(c.codec.front() == 0 && c.size() == 10 && c.encoder.count(0xf6) && c.encoder.count('~'));   // 0 ~ d e h l o r w ö
    CHECK(read_text32(gt) == utf8_to_utf32("# comment\nhello\n\nw\xc3\xb6rld"));
    vector<string> lines;
    read_lines(lines, gt);
    CHECK(lines.size() == 4 && lines[1] == "hello");
  }
  threw = false;
  try { make_Normalizer("fancy"); } catch (const char* msg) { threw = string(msg) == "unknown normalizer name"; }
  CHECK(threw);
  {
    std::unique_ptr<INormalizer> nm(make_Normalizer("center"));
    CHECK(nm->kind() == 2 && nm->target_height == 48 && nm->range == 4.0f && nm->smooth1d == 0.3f);
    std::unique_ptr<INormalizer> mm(make_Normalizer("mean"));
    CHECK(mm->kind() == 1 && mm->range == 1.0f);
  }
  // ---- PNG codec round trip (extras.cc:537-560): write quantises with floor(v*256), read returns (r+g+b)/(3*255.0)
  {
    Tensor2 im;
    im.resize(37, 11);
    for (int i = 0; i < 37; i++)
      for (int j = 0; j < 11; j++) im(i, j) = (float)((i * 7 + j * 13) % 101) / 100.0f;
    const string pn = string(fname) + ".png";
    write_png(pn.c_str(), im);
    Tensor2 back;
    read_png(back, pn.c_str());
    CHECK(back.dimension(0) == 37 && back.dimension(1) == 11);
    for (int i = 0; i < 37; i++)
      for (int j = 0; j < 11; j++) {
        double v = im(i, j) * 256;
        v = v > 255.999999 ? 255.999999 : v;
        const int q = (int)std::floor(v);
        CHECK(back(i, j) == (Float)((q + q + q) / (3 * 255.0)));
      }
    threw = false;
    try { read_png(back, "/nonexistent/x.png"); } catch (const char* msg) { threw = string(msg) == "error on open"; }
    CHECK(threw);
  }
  printf("host cpu ok: %d params saved to %s\n", P, fname);
  return 0;
}

// a RAW scanned-looking line (ink = 1): the bands of render() drawn 60 rows high with a margin, so that the
// CenterNormalizer has something to measure
static Tensor2 render_raw(const std::wstring& text) {
  Tensor2 small = render(text, 48), raw;
  const int T = small.dimension(0);
  raw.resize(T + 10, 60);
  for (int t = 0; t < T; t++)
    for (int i = 0; i < 48; i++) raw(t + 5, i + 7) = small(t, i);
  return raw;
}

static int run_gpu() {
  CLSTMOCR ocr;
  ocr.createBidi(demo_codec(), 16);
  ocr.normalizer.reset(make_NoNormalizer());   // render() draws lines that are already target_height rows high
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
  CHECK(ocr2.normalizer && ocr2.normalizer->kind() == 2);   // load installs a CenterNormalizer (clstmhl.h:165)
  ocr2.normalizer.reset(make_NoNormalizer());
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
  // ---- raw lines through PNG files and the CenterNormalizer (the clstmocrtrain data path, clstmocrtrain.cc:68-76)
  {
    CLSTMOCR o3;
    o3.createBidi(demo_codec(), 16);
    CHECK(o3.normalizer->kind() == 2);
    o3.setLearningRate(1e-2, 0.9);
    Tensor2 raw0 = render_raw(text), page, raw;
    page.resize(raw0.dimension(0), raw0.dimension(1));
    for (size_t k = 0; k < page.data.size(); k++) page.data[k] = 1.0f - raw0.data[k];   // paper is white on disk
    write_png("/tmp/clstm_b200_host_line.png", page);
    read_png(raw, "/tmp/clstm_b200_host_line.png");
    for (Float& v : raw.data) v = -v + Float(1);
    std::wstring g3;
    int it3 = 0;
    for (; it3 < 800; it3++) {
      g3 = o3.train(raw, text);
      if (g3 == text && it3 > 20) break;
    }
    printf("raw line: trained %d steps, reads: %s, normalised to %d x %d\n", it3, utf32_to_utf8(g3).c_str(),
           o3.image.dimension(0), o3.image.dimension(1));
    CHECK(o3.predict(raw) == text);
    CHECK(o3.image.dimension(1) == 48 && o3.image.dimension(0) > 10);
    // the stand-alone normalizer returns the very image the OCR wrapper fed to the network
    std::unique_ptr<INormalizer> nm(make_CenterNormalizer());
    Tensor2 nimg;
    nm->measure(raw);
    nm->normalize(nimg, raw);
    CHECK(nimg.dimension(0) == o3.image.dimension(0) && nimg.data == o3.image.data);
    bool threw = false;
    Tensor2 other = render_raw(L"ab");
    try { nm->normalize(nimg, other); } catch (const char* msg) { threw = string(msg) == "measure doesn't match normalize"; }
    CHECK(threw);
    vector<Tensor2> raws = {render_raw(L"ab"), render_raw(L"cde a"), raw};
    for (int k = 0; k < 200; k++) o3.train_batch(raws, tgs);
    auto r3 = o3.train_batch(raws, tgs);
    CHECK(r3[0] == tgs[0] && r3[1] == tgs[1] && r3[2] == tgs[2]);
    // two-deep input pipeline (what clstmocrtrain batch=N runs): prefetch the next batch while the step is in flight
    o3.prefetch_batch(raws, tgs);
    std::vector<std::wstring> rp;
    for (int k = 0; k < 30; k++) {
      o3.train_prefetched();
      o3.prefetch_batch(raws, tgs);
      rp = o3.fetch_results();
    }
    CHECK(rp.size() == 3 && rp[0] == tgs[0] && rp[1] == tgs[1] && rp[2] == tgs[2]);
    CHECK(o3.predict(raw) == text);                 // a synchronous call between pipeline steps leaves the prefetched batch intact
    o3.train_prefetched();
    rp = o3.fetch_results();
    CHECK(rp[2] == tgs[2]);
    bool threw2 = false;
    try { o3.train_prefetched(); } catch (const char*) { threw2 = true; }
    CHECK(threw2);
  }
  // ---- rank 4: a two-block net (bidi2) learns the same line; a sigmoid-output lstm1 runs forward/backward/update
  {
    CLSTMOCR o4;
    o4.net = make_net("bidi2", {{"ninput", 48}, {"nhidden", 12}, {"nhidden2", 10}, {"noutput", 7}});
    o4.net->codec.set(demo_codec());
    o4.nclasses = 7;
    o4.normalizer.reset(make_NoNormalizer());
    o4.setLearningRate(1e-2, 0.9);
    std::wstring g4;
    int it4 = 0;
    for (; it4 < 900; it4++) {
      g4 = o4.train(img, text);
      if (g4 == text && it4 > 20) break;
    }
    printf("bidi2: trained %d steps, reads: %s\n", it4, utf32_to_utf8(g4).c_str());
    CHECK(o4.predict(img) == text);
    o4.save("/tmp/clstm_b200_host_bidi2.clstm");
    CLSTMOCR o5;
    o5.load("/tmp/clstm_b200_host_bidi2.clstm");
    o5.normalizer.reset(make_NoNormalizer());
    CHECK(o5.predict(img) == text);
    Network l1 = make_net("lstm1", {{"ninput", 48}, {"nhidden", 6}, {"noutput", 1}, {"lstm_type", "RELUTANHNPLSTM"}});
    l1->setLearningRate(1e-3, 0.9);
    set_inputs(l1, img);
    l1->forward();
    CHECK(l1->outputs.size() == img.dimension(0) && l1->outputs.rows() == 1);
    for (int t = 0; t < l1->outputs.size(); t++) {
      CHECK(l1->outputs[t].v(0, 0) > 0.f && l1->outputs[t].v(0, 0) < 1.f);
      l1->outputs[t].d(0, 0) = 1.0f - l1->outputs[t].v(0, 0);          // push the sigmoid towards 1
    }
    const Float before = l1->outputs[10].v(0, 0);
    l1->backward();
    sgd_update(l1);
    for (int k = 0; k < 20; k++) {
      l1->forward();
      for (int t = 0; t < l1->outputs.size(); t++) l1->outputs[t].d(0, 0) = 1.0f - l1->outputs[t].v(0, 0);
      l1->backward();
      sgd_update(l1);
    }
    l1->forward();
    CHECK(l1->outputs[10].v(0, 0) > before);
    Network b0 = make_net("bidi0", {{"ninput", 48}, {"noutput", 3}});
    bool threw = false;
    set_inputs(b0, img);
    try { b0->forward(); } catch (const char*) { threw = true; }   // a bare Parallel is a container; it needs a Stacked root
    CHECK(threw);
  }
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

// decodes a PNG written by anybody (tests/golden/png, made with PIL) and prints r+g+b per pixel for cross-checking
static int run_png(const char* fname) {
  Tensor2 im;
  read_png(im, fname);
  printf("%d %d\n", im.dimension(0), im.dimension(1));
  for (int j = 0; j < im.dimension(1); j++) {
    for (int i = 0; i < im.dimension(0); i++) printf("%d ", (int)std::lround(im(i, j) * 3 * 255.0));
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && !strcmp(argv[1], "png")) return run_png(argv[2]);
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
