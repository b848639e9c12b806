// Command line of the drop-in executable: the reference's flag names, defaults and interactions
// (src/main.cpp:59-153; parser semantics from core/Parameters.cpp:533-565: short options take the
// next argument or an attached value, long options match by unique prefix and accept
// "--name value" or "--name:value", booleans toggle, unknown arguments are warned about and ignored).
#pragma once
#include <string>
#include <vector>

#include "polymutt_b200.h"

namespace pmh {

struct Options {
  std::string ped_file, dat_file, glf_index_file;
  std::string vcf_in, vcf_out, pos_file, chrs2process;
  std::string chrX = "X", chrY = "Y", chrMT = "MT";
  double posterior = 0.5;
  double theta = 0.001, theta_indel = 0.0001, tstv_ratio = 2.0, precision = 0.0001;
  int num_threads = 1;
  bool denovo = false;
  double denovo_mut_rate = 1.5e-08, denovo_tstv_ratio = 2.0, denovo_lr = 0.01;
  int min_map_quality = 0, min_total_depth = 0, max_total_depth = 0;
  double min_ps = 0;
  bool out_all_sites = false, gl_off = false, quick_call = false;
  // derived (main.cpp:151-153)
  bool force_call = false;
  // extensions of this implementation (not in the reference)
  int gpus = 1;            // --gpus N: shard consecutive site batches (--in_vcf: record ranges of a chunk) over N GPUs (devices device..device+N-1)
  int ingest_threads = 0;  // --ingest_threads N: GLF decode/merge threads (0 = all cores, capped at 32)
  int device = 0;          // --device
  int batch_sites = 0;     // --batch_sites (0 = automatic)
  std::string cmd;         // argv joined with spaces, trailing space (main.cpp:159-164)
  std::vector<std::string> warnings;

  // Returns false and sets *err on fatal problems (the reference's error() exits).
  bool parse(int argc, char **argv, std::string *err);
  void to_params(pm_params *p) const;
  void print_status() const;
};

}  // namespace pmh
