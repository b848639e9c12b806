#include "vcf_writer.h"

#include <ctime>

namespace pmh {

static const char kBases[5] = {'0', 'A', 'C', 'G', 'T'};
static const char *kGenoLabel[10] = {"A/A", "A/C", "A/G", "A/T", "C/C", "C/G", "C/T", "G/G", "G/T", "T/T"};

void VcfWriter::header(bool denovo) {
  time_t t;
  time(&t);
  bool af = ped_.families.size() > 1 || !ped_.families[0].nuclear();
  fprintf(fh_, "##fileformat=VCFv4.0\n");
  fprintf(fh_, "##fileDate=%s", ctime(&t));
  fprintf(fh_, "##command=%s\n", opt_.cmd.c_str());
  fprintf(fh_, "##minMapQuality=%f\n", (double)opt_.min_map_quality);
  fprintf(fh_, "##minTotalDepth=%d\n", opt_.min_total_depth);
  fprintf(fh_, "##maxTodalDepth=%d\n", opt_.max_total_depth);
  fprintf(fh_, "##posterior=%.3f\n", opt_.posterior);
  fprintf(fh_, "##INFO=<ID=NS,Number=1,Type=Integer,Description=\"Number of Samples With Data\">\n");
  fprintf(fh_, "##INFO=<ID=PS,Number=1,Type=Integer,Description=\"Percentage of Samples With Data\">\n");
  fprintf(fh_, "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total Read Depth\">\n");
  fprintf(fh_, "##INFO=<ID=MQ,Number=1,Type=Float,Description=\"Average Map Quality\">\n");
  if (af) fprintf(fh_, "##INFO=<ID=AF,Number=.,Type=Float,Description=\"Reference Allele Frequency\">\n");
  if (denovo) fprintf(fh_, "##INFO=<ID=DQ,Number=1,Type=Float,Description=\"De Novo Mutation Quality\">\n");
  fprintf(fh_, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
  fprintf(fh_, "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">\n");
  fprintf(fh_, "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read Depth\">\n");
  if (!denovo) fprintf(fh_, "##FORMAT=<ID=DS,Number=1,Type=Float,Description=\"Dosage: Defined As the Expected Alternative Allele Count\">\n");
  if (!opt_.gl_off) fprintf(fh_, "##FORMAT=<ID=PL,Number=10,Type=Integer,Description=\"Phred-scaled Genotype Likelhood\">\n");
  if (!denovo && opt_.force_call) fprintf(fh_, "##FORMAT=<ID=BA,String,Description=\"Best Alterantive Allele\">\n");
  fprintf(fh_, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT");
  for (int idx : ped_.columns()) fprintf(fh_, "\t%s", ped_.persons[idx].pid.c_str());
  fprintf(fh_, "\n");
  if (denovo) fflush(fh_);
  header_done_ = true;
}

void VcfWriter::write_site(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                           const pm_person_site *persons, const pm_person_result *pr) {
  if (opt_.denovo) write_denovo(chrom, hdr, r, persons, pr);
  else write_normal(chrom, hdr, r, persons, pr);
}

static inline int depth_of(const pm_person_site &p) { return p.depth[0] | (p.depth[1] << 8) | (p.depth[2] << 16); }

void VcfWriter::write_normal(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                             const pm_person_site *persons, const pm_person_result *pr) {
  if (!header_done_) header(false);
  const bool single_nuclear = ped_.families.size() == 1 && ped_.families[0].nuclear();
  const bool mono = (r.flags & PM_FLAG_MONO) != 0;
  const int a1 = r.allele1, a2 = r.allele2, ref = hdr.ref_base;
  char info[512];
  int n;
  if (single_nuclear)
    n = snprintf(info, sizeof info, "NS=%d;PS=%.1f;DP=%d;MQ=%.1f", r.num_samp, r.perc_samp * 100, r.total_depth, r.avg_map_qual);
  else if (hdr.chr_class != PM_CHR_AUTO)
    n = snprintf(info, sizeof info, "NS=%d;PS=%.1f;DP=%d;MQ=%.1f;AF=%.4f", r.num_samp, r.perc_samp * 100, r.total_depth, r.avg_map_qual, r.freq);
  else
    n = snprintf(info, sizeof info, "NS=%d;PS=%.1f;DP=%d;MQ=%.1f;AF=%.4f;AB=%.3f", r.num_samp, r.perc_samp * 100, r.total_depth, r.avg_map_qual, r.freq, r.ab);
  if (mono) snprintf(info + n, sizeof info - n, ";BA=%c", kBases[a2]);
  std::string alt;
  if (ref == a1) alt = std::string(1, kBases[mono ? a1 : a2]);
  else { alt += kBases[a1]; alt += ","; alt += kBases[a2]; }
  fprintf(fh_, "%s\t%d\t%s\t%c\t%s\t%d\t%s\t%s\t%s", chrom.c_str(), (int)hdr.pos + 1, ".", kBases[ref], alt.c_str(),
          int(r.poly_qual + 0.5), ".", info, opt_.gl_off ? "GT:GQ:DP:DS" : "GT:GQ:DP:DS:PL");
  static const char *lab[5] = {"0/0", "0/1", "1/1", "1/2", "2/2"};
  static const char *lab_hap[5] = {"0", "ERROR", "1", "ERROR2", "2"};
  const int g11 = genotype_index(a1, a1), g12 = genotype_index(a1, a2), g22 = genotype_index(a2, a2);
  const int np = ped_.n_person();
  for (int i = 0; i < np; i++) {
    int best = pr[i].best;
    int label_idx = (ref == a1) ? best : best + 2;
    const char *gt = lab[label_idx];
    if (hdr.chr_class == PM_CHR_Y || hdr.chr_class == PM_CHR_MT) gt = lab_hap[label_idx];
    fprintf(fh_, "\t%s:", gt);
    fprintf(fh_, "%d:", (int)pr[i].gq);
    fprintf(fh_, "%d:", depth_of(persons[i]));
    fprintf(fh_, "%.2f", pr[i].dosage);
    if (!opt_.gl_off) fprintf(fh_, ":%u,%u,%u", persons[i].lk[g11], persons[i].lk[g12], persons[i].lk[g22]);
  }
  fprintf(fh_, "\n");
  fflush(fh_);
  rows_++;
}

void VcfWriter::write_denovo(const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                             const pm_person_site *persons, const pm_person_result *pr) {
  if (!header_done_) header(true);
  const bool single_nuclear = ped_.families.size() == 1 && ped_.families[0].nuclear();
  const bool mono = (r.flags & PM_FLAG_MONO) != 0;
  const int a1 = r.allele1, ref = hdr.ref_base;
  const int a2_label = r.allele2;           // alleles the genotype labels were made with
  const int a2 = mono ? a1 : r.allele2;     // denovo_mono: allele2 = allele1 (NucFam.cpp:1870)
  char info[512];
  if (single_nuclear)
    snprintf(info, sizeof info, "NS=%d;PS=%.1f;DP=%d;MQ=%.1f;DQ=%.3f", r.num_samp, r.perc_samp * 100, r.total_depth, r.avg_map_qual, r.denovo_lr);
  else
    snprintf(info, sizeof info, "NS=%d;PS=%.1f;DP=%d;MQ=%.1f;AF=%.4f;DQ=%.3f", r.num_samp, r.perc_samp * 100, r.total_depth, r.avg_map_qual, r.freq, r.denovo_lr);
  std::string alt;
  if (ref == a1) alt = std::string(1, kBases[a2]);
  else { alt += kBases[a1]; alt += ","; alt += kBases[a2]; }
  fprintf(fh_, "%s\t%d\t%s\t%c\t%s\t%d\t%s\t%s\t%s", chrom.c_str(), (int)hdr.pos + 1, ".", kBases[ref], alt.c_str(),
          int(r.poly_qual + 0.5), ".", info, opt_.gl_off ? "GT:GQ:DP" : "GT:GQ:DP:PL");
  static const char *lab[5] = {"0/0", "0/1", "1/1", "1/2", "2/2"};
  int col = 0;
  for (const Family &f : ped_.families) {
    const bool letters = (int)f.path.size() != f.founders;  // nuclear / extended families print base letters
    for (size_t j = 0; j < f.path.size(); j++, col++) {
      const pm_person_result &p = pr[col];
      const char *gt;
      if (!letters) gt = lab[(ref == a1) ? p.best : p.best + 2];  // CalcPostProb_SinglePerson -> vcfv4 label
      else if (p.ten_state) gt = kGenoLabel[p.best];
      else {
        int idx = p.best == 0 ? genotype_index(a1, a1) : p.best == 1 ? genotype_index(a1, a2_label) : genotype_index(a2_label, a2_label);
        gt = kGenoLabel[idx];
      }
      fprintf(fh_, "\t%s:", gt);
      fprintf(fh_, "%d:", (int)p.gq);
      fprintf(fh_, "%d", depth_of(persons[col]));
      if (!opt_.gl_off) {
        fprintf(fh_, ":");
        for (int g = 0; g < 9; g++) fprintf(fh_, "%d,", persons[col].lk[g]);
        fprintf(fh_, "%d", persons[col].lk[9]);
      }
    }
  }
  fprintf(fh_, "\n");
  fflush(fh_);
  rows_++;
}

}  // namespace pmh
