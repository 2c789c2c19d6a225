// clstmocr -- drop-in for the reference's recognition CLI (clstmocr.cc:1-117) on top of the device library:
// reads a list of line images, prints "<file>\t<text>", optionally writes <base>.txt, character-wise predictions
// (conf=1) and posterior images (output=posteriors|logs).
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "clstm_extras.h"
#include "clstm_host.h"

using namespace ocropus;
using std::cerr;
using std::cout;
using std::endl;
using std::string;
using std::vector;

static float scaled_log(float x) {   // clstmocr.cc:33-40: log posterior mapped from [-10, 0] to [0, 1]
  const float thresh = 10.0;
  if (x <= 0.0) return 0.0;
  const float l = log(x);
  if (l < -thresh) return 0.0;
  if (l > 0) return 1.0;
  return (l + thresh) / thresh;
}

static int print_usage(char** argv) {
  cerr << "Usage: [VAR=VAL...] " << argv[0] << " IMAGEFILE\n\n"
       << "  Arguments:\n"
       << "    IMAGEFILE      Image file to OCR\n\n"
       << "  Variables:\n"
       << "     load          Model to recognize with. Required\n"
       << "     conf          Output character-wise predictions. Default: 0\n"
       << "     output        Output format, either 'text' or 'posteriors'. Default: 'text'\n"
       << "     save_text     Save text to IMAGEFILE.txt. Default: 1\n";
  return EXIT_FAILURE;
}

static int main1(int argc, char** argv) {
  if (argc != 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) return print_usage(argv);
  const string load_name = getsenv("load", "");
  if (load_name == "") THROW("must give load= parameter");
  CLSTMOCR clstm;
  clstm.load(load_name);

  const bool conf = getienv("conf", 0);
  const string output = getsenv("output", "text");
  const bool save_text = getienv("save_text", 1);

  std::ifstream stream(argv[1]);   // the argument is a LIST of image files, one per line (clstmocr.cc:73-75)
  string line;
  while (getline(stream, line)) {
    Tensor2 raw;
    const string base = line.substr(0, line.find_last_of("."));
    read_png(raw, line.c_str());
    for (Float& v : raw.data) v = -v + Float(1.0);
    if (!conf) {
      const string out = clstm.predict_utf8(raw);
      cout << line << "\t" << out << endl;
      if (save_text) write_text(base + ".txt", out);
    } else {
      cout << "file " << line << endl;
      vector<CharPrediction> preds;
      clstm.predict(preds, raw);
      for (const CharPrediction& p : preds) cout << p.i << "\t" << p.x << "\t" << p.c << "\t" << p.p << endl;
    }
    if (output == "text") {
    } else if (output == "logs" || output == "posteriors") {
      Tensor2 outputs;
      clstm.get_outputs(outputs);
      if (output == "logs")
        for (Float& v : outputs.data) v = scaled_log(v);
      write_png((base + (output == "logs" ? ".lp.png" : ".p.png")).c_str(), outputs);
    } else {
      THROW("unknown output format");
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  try {
    return main1(argc, argv);
  } catch (const char* message) {
    cerr << "FATAL: " << message << endl;
  } catch (const std::string& message) {
    cerr << "FATAL: " << message << endl;
  }
  return 1;
}
