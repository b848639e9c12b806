// VCF text writers for GLF-input runs: byte-for-byte the output of
// NucFamGenotypeLikelihood::OutputVCF (src/NucFamGenotypeLikelihood.cpp:1751-1830) and
// OutputVCF_denovo (:1832-1915), fed from pm_site_result / pm_person_result instead of class members.
#pragma once
#include <cstdio>
#include <string>

#include "params.h"
#include "pedigree.h"
#include "polymutt_b200.h"

namespace pmh {

// printf("%.<decimals>f") restated: appends exactly what glibc prints (round-half-even on the exact binary
// value, "-0.00" for negative zeros), ~20x faster.  Non-finite or huge values fall back to snprintf.
void append_fixed(std::string &out, double x, int decimals);
void append_int(std::string &out, long long v);

class VcfWriter {
 public:
  VcfWriter(FILE *fh, const Options &opt, const Pedigree &ped) : fh_(fh), opt_(opt), ped_(ped) {}
  // One emitted site. `persons` = the site's packed input records, `pr` = its per-person results.
  void write_site(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                  const pm_person_site *persons, const pm_person_result *pr);
  // The row's text appended to `out` (no header, no I/O, no state: safe to call from several threads at once,
  // which is how the driver formats a batch of rows off the consumer thread).
  void format_site(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                   const pm_person_site *persons, const pm_person_result *pr) const;
  // Rows formatted by format_site, in order: prints the header first if it has not been printed.
  void write_rows(const char *text, size_t bytes, long n_rows);
  // OutputVCF_denovo prints the header the first time it is entered, even if it then drops the row
  // (NucFam.cpp:1834-1868): called when a PM_SITE_DENOVO_DROPPED site is seen.
  void ensure_header() { if (!header_done_) header(opt_.denovo); }
  long rows_written() const { return rows_; }

 private:
  void header(bool denovo);
  void format_normal(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                     const pm_person_site *persons, const pm_person_result *pr) const;
  void format_denovo(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                     const pm_person_site *persons, const pm_person_result *pr) const;
  FILE *fh_;
  const Options &opt_;
  const Pedigree &ped_;
  bool header_done_ = false;
  long rows_ = 0;
};

inline int genotype_index(int b1, int b2) {  // core/glfHandler.h:102-106
  return b1 < b2 ? (b1 - 1) * (10 - b1) / 2 + (b2 - b1) : (b2 - 1) * (10 - b2) / 2 + (b1 - b2);
}

}  // namespace pmh
