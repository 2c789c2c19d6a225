// clstmocrtrain -- drop-in for the reference's training CLI (clstmocrtrain.cc:1-224) on top of the device library.
// Same command line, environment variables, sampling (lrand48), triggers, log lines and checkpoint names.
// One extension: batch=N (default 1) draws N samples per trial and runs them as ONE device step
// (normalise + forward + CTC + backward + update, CLSTMOCR::train_batch); batch=1 is the reference's per-line loop.
// The display server (display_every, PyServer) is out of scope.
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "clstm_extras.h"
#include "clstm_host.h"

using namespace ocropus;
using std::cerr;
using std::string;
using std::vector;
using std::wstring;

static wstring separate_chars(const wstring& s, const wstring& charsep) {   // clstmocrtrain.cc:46-54
  if (charsep.empty()) return s;
  wstring result;
  for (size_t i = 0; i < s.size(); i++) {
    if (i > 0) result.push_back(charsep[0]);
    result.push_back(s[i]);
  }
  return result;
}

struct Dataset {   // clstmocrtrain.cc:56-77
  vector<string> fnames;
  wstring charsep = utf8_to_utf32(getsenv("charsep", ""));
  int size() { return (int)fnames.size(); }
  Dataset() {}
  Dataset(string file_list) { readFileList(file_list); }
  void readFileList(string file_list) { read_lines(fnames, file_list); }
  void getCodec(Codec& codec) {
    vector<string> gtnames;
    for (auto s : fnames) gtnames.push_back(basename(s) + ".gt.txt");
    codec.build(gtnames, charsep);
  }
  void readSample(Tensor2& raw, wstring& gt, int index) {
    const string fname = fnames[index];
    gt = separate_chars(read_text32(basename(fname) + ".gt.txt"), charsep);
    read_png(raw, fname.c_str());
    for (Float& v : raw.data) v = -v + Float(1);   // ink = 1
  }
};

static std::pair<double, double> test_set_error(CLSTMOCR& clstm, Dataset& testset) {   // clstmocrtrain.cc:79-91
  double count = 0.0, errors = 0.0;
  for (int test = 0; test < testset.size(); test++) {
    Tensor2 raw;
    wstring gt;
    testset.readSample(raw, gt, test);
    wstring pred = clstm.predict(raw);
    count += gt.size();
    errors += levenshtein(pred, gt);
  }
  return std::make_pair(errors, count);
}

static int print_usage(char** argv) {
  cerr << "Usage: [VAR=VAL...] " << argv[0] << " TRAININGLIST [TESTLIST]\n\n"
       << "  Arguments:\n"
       << "    TRAININGLIST     File with filenames to train with\n"
       << "    TESTLIST         File with filenames to evaluate training\n\n"
       << "  Variables:\n"
       << "     load            Filename of model file to load. Default: ''\n"
       << "     save_name       Basename of model file to save. Default: '_ocr'\n"
       << "     nhidden         Number of hidden Default: 100\n"
       << "     lrate           Learning rate. Default: 1e-4\n"
       << "     momentum        Momentum. Default: 0.9\n"
       << "     target_height   Line height to normalize. Default: 48\n"
       << "     ntrain          Number of iterations. Default: 10000000\n"
       << "     start           Initial iteration. Default: -1\n"
       << "     charsep         Separator between characters in ground truth. Default: ''\n"
       << "     report_time     Set to 1 to report time. Default: 0\n"
       << "     test_every      Evaluate model every n-th iteration. Default: 10000\n"
       << "     report_every    Log current state every n-th iteration. Default: 100\n"
       << "     save_every      Save model with iteration as suffix every n-th\n"
       << "                     iteration. Default: 10000\n"
       << "     batch           Lines per device step (extension). Default: 1\n"
       << "     params          Whether to report variable values on read. Default: 1\n";
  return EXIT_FAILURE;
}

static int main1(int argc, char** argv) {
  if (argc < 2 || argc > 3 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) return print_usage(argv);
  const int ntrain = getienv("ntrain", 10000000);
  const string save_name = getsenv("save_name", "_ocr");
  const int report_time = getienv("report_time", 0);
  const int batch = std::max(1, getienv("batch", 1));

  Dataset trainingset(argv[1]);
  if (trainingset.size() <= 0) THROW("empty training list");
  Dataset testset;
  if (argc > 2) testset.readFileList(argv[2]);
  print("got", trainingset.size(), "files,", testset.size(), "tests");

  const string load_name = getsenv("load", "");
  CLSTMOCR clstm;
  if (load_name != "") {
    clstm.load(load_name);
  } else {
    Codec codec;
    trainingset.getCodec(codec);
    print("got", codec.size(), "classes");
    clstm.target_height = int(getrenv("target_height", 48));
    clstm.createBidi(codec.codec, getienv("nhidden", 100));
    clstm.setLearningRate(getdenv("lrate", 1e-4), getdenv("momentum", 0.9));
  }
  network_info(clstm.net, "");

  double test_error = 9999.0;
  double best_error = 1e38;
  double start_time = now();
  const int start = (int)(double)clstm.net->attr.get("trial", getienv("start", -1)) + 1;
  if (start > 0) print("start", start);

  Trigger test_trigger(getienv("test_every", 10000), -1, start);
  test_trigger.skip0();
  Trigger save_trigger(getienv("save_every", 10000), ntrain, start);
  save_trigger.enable(save_name != "").skip0();
  Trigger report_trigger(getienv("report_every", 100), ntrain, start);

  // batch > 1: two-deep input pipeline -- while the device runs step `trial`, the host reads and the copy stream
  // normalises the lines of step `trial + 1`
  vector<Tensor2> raws(batch);
  vector<wstring> gts(batch), next_gts(batch);
  auto draw = [&](vector<wstring>& into) {
    for (int k = 0; k < batch; k++) trainingset.readSample(raws[k], into[k], lrand48() % trainingset.size());
    clstm.prefetch_batch(raws, into);
  };
  if (batch > 1 && start < ntrain) draw(next_gts);
  for (int trial = start; trial < ntrain; trial++) {
    wstring gt, pred;
    if (batch == 1) {
      const int sample = lrand48() % trainingset.size();
      Tensor2 raw;
      trainingset.readSample(raw, gt, sample);
      pred = clstm.train(raw, gt);
    } else {
      gts.swap(next_gts);
      clstm.train_prefetched();                       // launches the step on the prefetched batch and returns
      if (trial + 1 < ntrain) draw(next_gts);         // host PNG decoding + device normalisation overlap the step
      vector<wstring> preds = clstm.fetch_results();
      gt = gts[0];
      pred = preds[0];
    }

    if (report_trigger(trial)) {
      print(trial);
      print("TRU", gt);
      if (batch == 1) print("ALN", clstm.aligned_utf8());
      print("OUT", utf32_to_utf8(pred));
      if (trial > 0 && report_time) print("steptime", (now() - start_time) / report_trigger.since());
      start_time = now();
    }

    if (test_trigger(trial)) {
      auto tse = test_set_error(clstm, testset);
      const double errors = tse.first, count = tse.second;
      test_error = errors / count;
      print("ERROR", trial, test_error, "   ", errors, count);
      if (test_error < best_error) {
        best_error = test_error;
        const string fname = save_name + ".clstm";
        print("saving best performing network so far", fname, "error rate: ", best_error);
        clstm.net->attr.set("trial", trial);
        clstm.save(fname);
      }
    }

    if (save_trigger(trial)) {
      const string fname = save_name + "-" + std::to_string(trial) + ".clstm";
      print("saving", fname);
      clstm.net->attr.set("trial", trial);
      clstm.save(fname);
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  try {
    return main1(argc, argv);
  } catch (const char* message) {
    cerr << "FATAL: " << message << std::endl;
  } catch (const std::string& message) {
    cerr << "FATAL: " << message << std::endl;
  }
  return 1;
}
