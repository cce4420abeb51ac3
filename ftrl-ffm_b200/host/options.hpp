// options.hpp -- command line of the drop-in `main`.
// Same flags, defaults and input-type sniffing as the reference: src/utils/cmd_option.cpp:61-113,
// src/include/utils/cmd_option.h:7-63.  Additive flags (default = reference-compatible): --batch_size,
// --device, --n_gpus, --seed, --csr_cache, --auc.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace host {

static const char *kHelp =
    "\nUsage: ./main [<options>]"
    "\n"
    "\n"
    "options:\n"
    "--model_path <model_path>: set the output model path\n"
    "--train_data <data_path>: set the train data path\n"
    "--eval_data <data_path>: set the eval data path\n"
    "--model_type <model_type>: LR, FM or FFM\n"
    "--init_mean <mean>: mean for parameter initialization\tdefault:0.0\n"
    "--init_stddev <stddev>: stddev for parameter initialization\tdefault:0.02\n"
    "--n_fields <n_fields>: number of fields in FFM\tdefault:8\n"
    "--n_feats <n_feats>: number of total features\tdefault:10000\n"
    "--n_factors <n_factors>: number of embed size in FM and FFM\tdefault:16\n"
    "--w_alpha <w_alpha>: alpha is one of the learning rate parameters\tdefault:1e-4\n"
    "--w_beta <w_beta>: beta is one of the learning rate parameters\tdefault:1.0\n"
    "--w_l1 <w_L1_reg>: L1 regularization parameter of w\tdefault:0.1\n"
    "--w_l2 <w_L2_reg>: L2 regularization parameter of w\tdefault:5.0\n"
    "--n_threads <threads_num>: set the number of (parser) threads\tdefault:1\n"
    "--n_epochs <epochs>: how many epochs to train\tdefault:1\n"
    "--online <online>: whether to online training mode\tdefault:true\n"
    "--batch_size <n>: samples per GPU minibatch; 1 = reference-exact sequential mode\tdefault:1024\n"
    "--device <ordinal>: CUDA device (first device of a multi-GPU run)\tdefault:0\n"
    "--n_gpus <n>: 1, 2, 4 or 8 GPUs of this box; the tables are sharded by feature id, every minibatch is split "
    "across the GPUs (data in memory: --online false or --csr_cache true)\tdefault:1\n"
    "--seed <n>: shuffle / init seed (0 = from std::random_device)\tdefault:0\n"
    "--csr_cache <bool>: keep a binary image <file>.csr of each parsed data file and reuse it while the text is "
    "unchanged\tdefault:false\n"
    "--auc <bool>: also print `epoch N eval auc: ...` after every evaluation pass (computed on the GPU)\tdefault:false\n";

struct Options {  // mirrors config_options (cmd_option.h:29-63)
  std::string model_path, train_path, eval_path, model_type = "FFM", file_type;
  float init_mean = 0.0f, init_stddev = 0.02f, w_alpha = 1e-4f, w_beta = 1.0f, w_l1 = 0.1f, w_l2 = 5.0f;
  int thread_num = 1, epoch = 1, n_fields = 8, n_feats = 10000, n_factors = 16;
  bool cmd = false, online = true;
  // additive
  long batch_size = 1024;
  bool csr_cache = false;
  bool auc = false;
  int device = 0;
  int n_gpus = 1;
  unsigned long long seed = 0;
};

inline bool assign_bool(std::string a) {  // cmd_option.cpp:24-27
  std::transform(a.begin(), a.end(), a.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return a == "true" || a == "1";
}
inline std::string upper(std::string a) {  // cmd_option.cpp:29-33
  std::transform(a.begin(), a.end(), a.begin(), [](unsigned char c) { return (char)std::toupper(c); });
  return a;
}

// cmd_option.cpp:35-59: count ':' in the first feature token of the first line
inline std::string detect_file_type(const std::string &path) {
  std::ifstream ifs(path);
  if (!ifs.good()) {
    fprintf(stderr, "fail to open %s\n", path.c_str());
    exit(EXIT_FAILURE);
  }
  std::string line;
  std::getline(ifs, line);
  std::vector<std::string> toks;
  size_t b = line.find_first_not_of(' ');
  while (b != std::string::npos) {
    const size_t e = line.find_first_of(' ', b);
    toks.push_back(line.substr(b, e == std::string::npos ? std::string::npos : e - b));
    b = e == std::string::npos ? e : line.find_first_not_of(' ', e);
  }
  if (toks.size() < 2) {
    fprintf(stderr, "unknown file format...\n");
    exit(EXIT_FAILURE);
  }
  const long colons = std::count(toks[1].begin(), toks[1].end(), ':');
  if (colons == 1) return "libsvm";
  if (colons == 2) return "libffm";
  fprintf(stderr, "unknown file format...\n");
  exit(EXIT_FAILURE);
}

// cmd_option.cpp:61-113.  Unknown flag -> std::invalid_argument (main prints help and exits 1).
inline void parse_options(int argc, char **argv, Options &o) {
  std::vector<std::string> a(argv + 1, argv + argc);
  if (a.size() % 2 != 0) throw std::invalid_argument("every --flag needs a value\n");  // reference: assert
  for (size_t i = 0; i < a.size(); i += 2) {
    const std::string &k = a[i], &v = a[i + 1];
    if (k == "--model_path") o.model_path = v;
    else if (k == "--model_type") o.model_type = upper(v);
    else if (k == "--online") o.online = assign_bool(v);
    else if (k == "--n_fields") o.n_fields = std::stoi(v);
    else if (k == "--n_feats") o.n_feats = std::stoi(v);
    else if (k == "--n_factors") o.n_factors = std::stoi(v);
    else if (k == "--train_data") o.train_path = v;
    else if (k == "--eval_data") o.eval_path = v;
    else if (k == "--init_mean") o.init_mean = std::stof(v);
    else if (k == "--init_stddev") o.init_stddev = std::stof(v);
    else if (k == "--w_alpha") o.w_alpha = std::stof(v);
    else if (k == "--w_beta") o.w_beta = std::stof(v);
    else if (k == "--w_l1") o.w_l1 = std::stof(v);
    else if (k == "--w_l2") o.w_l2 = std::stof(v);
    else if (k == "--n_threads") o.thread_num = std::stoi(v);
    else if (k == "--n_epochs") o.epoch = std::stoi(v);
    else if (k == "--cmd") o.cmd = assign_bool(v);
    else if (k == "--batch_size") o.batch_size = std::stol(v);
    else if (k == "--device") o.device = std::stoi(v);
    else if (k == "--n_gpus") o.n_gpus = std::stoi(v);
    else if (k == "--seed") o.seed = std::stoull(v);
    else if (k == "--csr_cache") o.csr_cache = assign_bool(v);
    else if (k == "--auc") o.auc = assign_bool(v);
    else throw std::invalid_argument("unknown argument: " + k + "\n");
  }
  o.file_type = detect_file_type(o.train_path);
  if (o.model_type == "FFM" && o.file_type != "libffm") {  // cmd_option.cpp:109-113
    fprintf(stderr, "FFM model requires libffm data format...\n");
    exit(EXIT_FAILURE);
  }
  if (o.thread_num < 1) o.thread_num = 1;
  if (o.batch_size < 1) o.batch_size = 1;
}

}  // namespace host
