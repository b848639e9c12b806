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

class VcfWriter {
 public:
  VcfWriter(FILE *fh, const Options &opt, const Pedigree &ped) : fh_(fh), opt_(opt), ped_(ped) {}
  // One emitted site. `persons` = the site's packed input records, `pr` = its per-person results.
  void write_site(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                  const pm_person_site *persons, const pm_person_result *pr);
  // OutputVCF_denovo prints the header the first time it is entered, even if it then drops the row
  // (NucFam.cpp:1834-1868): called when a PM_SITE_DENOVO_DROPPED site is seen.
  void ensure_header() { if (!header_done_) header(opt_.denovo); }
  long rows_written() const { return rows_; }

 private:
  void header(bool denovo);
  void write_normal(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                    const pm_person_site *persons, const pm_person_result *pr);
  void write_denovo(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                    const pm_person_site *persons, const pm_person_result *pr);
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
