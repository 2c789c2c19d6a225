// clstmocrtrain -- training CLI on top of the device library, a drop-in for the reference tool of the same name
// (command line, environment knobs, lrand48 sampling, trigger schedule, log lines and checkpoint names follow
// /root/reference/clstmocrtrain.cc:93-224; behaviour is cited per block below).
// One extension: batch=N (default 1).  N > 1 draws N samples per trial and runs them as ONE device step through the
// two-deep input pipeline (CLSTMOCR::prefetch_batch / train_prefetched / fetch_results): while the device runs trial t the
// host decodes the PNGs of trial t+1 and the copy stream normalises them.  batch=1 is the reference's per-line loop.
// The display server (display_every, PyServer) is out of scope.
#include <algorithm>
#include <cstring>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "clstm_extras.h"
#include "clstm_host.h"

namespace {
using namespace ocropus;
using std::string;
using std::vector;
using std::wstring;

// ---- knobs (all from the environment, reported on first read like the reference's getenv helpers) -------------------
struct Knob { const char* name; const char* what; };
const Knob kKnobs[] = {
    {"load", "Filename of model file to load. Default: ''"},
    {"save_name", "Basename of model file to save. Default: '_ocr'"},
    {"nhidden", "Number of hidden Default: 100"},
    {"lrate", "Learning rate. Default: 1e-4"},
    {"momentum", "Momentum. Default: 0.9"},
    {"target_height", "Line height to normalize. Default: 48"},
    {"ntrain", "Number of iterations. Default: 10000000"},
    {"start", "Initial iteration. Default: -1"},
    {"charsep", "Separator between characters in ground truth. Default: ''"},
    {"report_time", "Set to 1 to report time. Default: 0"},
    {"test_every", "Evaluate model every n-th iteration. Default: 10000"},
    {"report_every", "Log current state every n-th iteration. Default: 100"},
    {"save_every", "Save model with iteration as suffix every n-th iteration. Default: 10000"},
    {"batch", "Lines per device step, two-deep input pipeline (extension). Default: 1"},
    {"params", "Whether to report variable values on read. Default: 1"},
};
int usage(const char* prog) {
  std::cerr << "Usage: [VAR=VAL...] " << prog << " TRAININGLIST [TESTLIST]\n\n  Arguments:\n"
            << "    TRAININGLIST     File with filenames to train with\n"
            << "    TESTLIST         File with filenames to evaluate training\n\n  Variables:\n";
  for (const Knob& k : kKnobs)
    std::cerr << "     " << k.name << string(std::max<size_t>(1, 16 - strlen(k.name)), ' ') << k.what << "\n";
  return EXIT_FAILURE;
}

// ---- a list of line images with their transcripts "<base>.gt.txt" (clstmocrtrain.cc:56-77) --------------------------
class LineList {
 public:
  explicit LineList(const wstring& sep) : sep_(sep) {}
  void load(const string& list_file) { read_lines(files_, list_file); }
  int count() const { return (int)files_.size(); }
  Codec codec() const {                               // every character of every transcript, plus the separator
    vector<string> transcripts;
    for (const string& f : files_) transcripts.push_back(basename(f) + ".gt.txt");
    Codec c;
    c.build(transcripts, sep_);
    return c;
  }
  // image with ink = 1 (the files hold dark ink on white paper) and the transcript, separator interleaved if any
  void sample(int index, Tensor2& ink, wstring& text) const {
    const string& f = files_[index];
    const wstring plain = read_text32(basename(f) + ".gt.txt");
    text.clear();
    for (size_t i = 0; i < plain.size(); i++) {
      if (i && !sep_.empty()) text.push_back(sep_[0]);
      text.push_back(plain[i]);
    }
    read_png(ink, f.c_str());
    for (Float& v : ink.data) v = Float(1) - v;
  }
  int random_index() const { return (int)(lrand48() % count()); }   // the reference's sampler

 private:
  vector<string> files_;
  wstring sep_;
};

// ---- the run --------------------------------------------------------------------------------------------------------
class Run {
 public:
  Run(const string& train_list, const char* test_list)
      : sep_(utf8_to_utf32(getsenv("charsep", ""))), train_(sep_), test_(sep_) {
    ntrain_ = getienv("ntrain", 10000000);
    save_name_ = getsenv("save_name", "_ocr");
    report_time_ = getienv("report_time", 0) != 0;
    batch_ = std::max(1, getienv("batch", 1));
    train_.load(train_list);
    if (train_.count() <= 0) THROW("empty training list");
    if (test_list) test_.load(test_list);
    print("got", train_.count(), "files,", test_.count(), "tests");
    make_model();
    network_info(ocr_.net, "");
    // a loaded checkpoint carries the trial it was written at; training resumes after it (clstmocrtrain.cc:146)
    first_ = (int)(double)ocr_.net->attr.get("trial", getienv("start", -1)) + 1;
    if (first_ > 0) print("start", first_);
  }

  void loop() {
    Trigger evaluate(getienv("test_every", 10000), -1, first_);
    evaluate.skip0();
    Trigger numbered(getienv("save_every", 10000), ntrain_, first_);
    numbered.enable(!save_name_.empty()).skip0();
    Trigger report(getienv("report_every", 100), ntrain_, first_);
    double tick = now(), best = 1e38;

    vector<Tensor2> inks(batch_);
    vector<wstring> texts(batch_), queued(batch_);
    if (batch_ > 1 && first_ < ntrain_) enqueue(inks, queued);
    for (int trial = first_; trial < ntrain_; trial++) {
      wstring truth, read;
      if (batch_ == 1) {                              // the reference's loop body: one line, one update
        Tensor2 ink;
        train_.sample(train_.random_index(), ink, truth);
        read = ocr_.train(ink, truth);
      } else {
        texts.swap(queued);
        ocr_.train_prefetched();                      // launches the step on the queued batch and returns
        if (trial + 1 < ntrain_) enqueue(inks, queued);   // PNG decoding + device normalisation overlap that step
        read = ocr_.fetch_results()[0];
        truth = texts[0];
      }
      if (report(trial)) {                            // log lines of clstmocrtrain.cc:181-189
        print(trial);
        print("TRU", truth);
        if (batch_ == 1) print("ALN", ocr_.aligned_utf8());
        print("OUT", utf32_to_utf8(read));
        if (trial > 0 && report_time_) print("steptime", (now() - tick) / report.since());
        tick = now();
      }
      if (evaluate(trial)) {
        const std::pair<double, double> e = test_errors();
        const double rate = e.first / e.second;
        print("ERROR", trial, rate, "   ", e.first, e.second);
        if (rate < best) {
          best = rate;
          print("saving best performing network so far", save_name_ + ".clstm", "error rate: ", best);
          checkpoint(save_name_ + ".clstm", trial);
        }
      }
      if (numbered(trial)) {
        const string f = save_name_ + "-" + std::to_string(trial) + ".clstm";
        print("saving", f);
        checkpoint(f, trial);
      }
    }
  }

 private:
  void make_model() {
    const string from = getsenv("load", "");
    if (!from.empty()) { ocr_.load(from); return; }
    const Codec c = train_.codec();
    print("got", c.size(), "classes");
    ocr_.target_height = int(getrenv("target_height", 48));
    ocr_.createBidi(c.codec, getienv("nhidden", 100));
    ocr_.setLearningRate(getdenv("lrate", 1e-4), getdenv("momentum", 0.9));
  }
  void enqueue(vector<Tensor2>& inks, vector<wstring>& texts) {
    for (int k = 0; k < batch_; k++) train_.sample(train_.random_index(), inks[k], texts[k]);
    ocr_.prefetch_batch(inks, texts);
  }
  std::pair<double, double> test_errors() {          // summed edit distance / summed transcript length (:79-91)
    double wrong = 0, total = 0;
    Tensor2 ink;
    wstring truth;
    for (int i = 0; i < test_.count(); i++) {
      test_.sample(i, ink, truth);
      wstring read = ocr_.predict(ink);
      wrong += levenshtein(read, truth);
      total += truth.size();
    }
    return {wrong, total};
  }
  void checkpoint(const string& file, int trial) {
    ocr_.net->attr.set("trial", trial);
    ocr_.save(file);
  }

  wstring sep_;
  LineList train_, test_;
  CLSTMOCR ocr_;
  string save_name_;
  int ntrain_ = 0, batch_ = 1, first_ = 0;
  bool report_time_ = false;
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2 || argc > 3 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) return usage(argv[0]);
  try {
    Run run(argv[1], argc > 2 ? argv[2] : nullptr);
    run.loop();
    return 0;
  } catch (const char* message) {                     // the reference's THROW convention
    std::cerr << "FATAL: " << message << std::endl;
  } catch (const std::string& message) {
    std::cerr << "FATAL: " << message << std::endl;
  }
  return 1;
}
