// clstmocr -- recognition CLI on top of the device library, a drop-in for the reference tool of the same name
// (/root/reference/clstmocr.cc:42-117): the argument is a LIST of line images; every line of output is "<file>\t<text>",
// the text also goes to <base>.txt (save_text), conf=1 prints per-character predictions, output=posteriors|logs writes
// the network outputs as an image next to the input.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "clstm_extras.h"
#include "clstm_host.h"

namespace {
using namespace ocropus;
using std::string;

int usage(const char* prog) {
  static const char* const knobs[][2] = {{"load", "Model to recognize with. Required"},
                                         {"conf", "Output character-wise predictions. Default: 0"},
                                         {"output", "Output format, 'text', 'posteriors' or 'logs'. Default: 'text'"},
                                         {"save_text", "Save text to IMAGEFILE.txt. Default: 1"}};
  std::cerr << "Usage: [VAR=VAL...] " << prog << " IMAGEFILE\n\n  Arguments:\n    IMAGEFILE      Image file to OCR\n\n  Variables:\n";
  for (auto& k : knobs) std::cerr << "     " << k[0] << string(std::max<size_t>(1, 14 - strlen(k[0])), ' ') << k[1] << "\n";
  return EXIT_FAILURE;
}

// log posterior mapped from [-10, 0] onto [0, 1] (clstmocr.cc:33-40)
float log_grey(float p) {
  if (p <= 0.0f) return 0.0f;
  const float l = std::log(p);
  return l < -10.0f ? 0.0f : (l > 0.0f ? 1.0f : (l + 10.0f) / 10.0f);
}

enum class Dump { none, posteriors, logs };

void recognise_list(const char* list_file) {
  const string model = getsenv("load", "");
  if (model.empty()) THROW("must give load= parameter");
  CLSTMOCR ocr;
  ocr.load(model);
  const bool per_char = getienv("conf", 0) != 0;
  const string fmt = getsenv("output", "text");
  const bool keep_text = getienv("save_text", 1) != 0;
  Dump dump = Dump::none;
  if (fmt == "posteriors") dump = Dump::posteriors;
  else if (fmt == "logs") dump = Dump::logs;
  else if (fmt != "text") THROW("unknown output format");

  std::ifstream list(list_file);
  for (string image; std::getline(list, image);) {
    const string stem = image.substr(0, image.find_last_of("."));
    Tensor2 ink;
    read_png(ink, image.c_str());
    for (Float& v : ink.data) v = Float(1) - v;       // ink = 1
    if (per_char) {
      std::cout << "file " << image << std::endl;
      std::vector<CharPrediction> chars;
      ocr.predict(chars, ink);
      for (const CharPrediction& c : chars) std::cout << c.i << "\t" << c.x << "\t" << c.c << "\t" << c.p << std::endl;
    } else {
      const string text = ocr.predict_utf8(ink);
      std::cout << image << "\t" << text << std::endl;
      if (keep_text) write_text(stem + ".txt", text);
    }
    if (dump != Dump::none) {
      Tensor2 post;
      ocr.get_outputs(post);
      if (dump == Dump::logs) std::transform(post.data.begin(), post.data.end(), post.data.begin(), log_grey);
      write_png((stem + (dump == Dump::logs ? ".lp.png" : ".p.png")).c_str(), post);
    }
  }
}
}  // namespace

int main(int argc, char** argv) {
  if (argc != 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) return usage(argv[0]);
  try {
    recognise_list(argv[1]);
    return 0;
  } catch (const char* message) {
    std::cerr << "FATAL: " << message << std::endl;
  } catch (const std::string& message) {
    std::cerr << "FATAL: " << message << std::endl;
  }
  return 1;
}
