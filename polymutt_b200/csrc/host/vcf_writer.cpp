#include "vcf_writer.h"

#include <cmath>
#include <cstring>
#include <ctime>

namespace pmh {

namespace {

// Raw-pointer cores: the caller guarantees room (kMaxInt / kMaxFixed bytes).
constexpr size_t kMaxInt = 24, kMaxFixed = 400;

inline char *put_int(char *o, long long v) {
  char buf[24];
  char *p = buf + sizeof buf;
  const bool neg = v < 0;
  unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
  do { *--p = (char)('0' + u % 10); u /= 10; } while (u);
  if (neg) *--p = '-';
  const size_t n = (size_t)(buf + sizeof buf - p);
  memcpy(o, p, n);
  return o + n;
}

struct U8Table {  // "0".."255"
  char txt[256][4];
  uint8_t len[256];
  U8Table() {
    for (int v = 0; v < 256; v++) len[v] = (uint8_t)snprintf(txt[v], 4, "%d", v);
  }
};
const U8Table kU8;
inline char *put_u8(char *o, unsigned v) {
  memcpy(o, kU8.txt[v], 4);
  return o + kU8.len[v];
}

char *put_fixed(char *o, double x, int decimals);

}  // namespace

void append_int(std::string &out, long long v) {
  char buf[kMaxInt];
  out.append(buf, (size_t)(put_int(buf, v) - buf));
}

void append_fixed(std::string &out, double x, int decimals) {
  char buf[kMaxFixed];
  out.append(buf, (size_t)(put_fixed(buf, x, decimals) - buf));
}

namespace {
char *put_fixed(char *o, double x, int decimals) {
  static const double kPow[7] = {1, 10, 100, 1000, 10000, 100000, 1000000};
  static const unsigned long long kPowI[7] = {1, 10, 100, 1000, 10000, 100000, 1000000};
  const double s = std::fabs(x);
  if (decimals < 0 || decimals > 6 || !(s < 1e9)) {  // also NaN / inf
    return o + snprintf(o, kMaxFixed, "%.*f", decimals, x);
  }
  // s * 10^d = p + e exactly (p the rounded product, e the fma residual); the digit string is the exact value
  // rounded to an integer, ties to even
  const double p = s * kPow[decimals];
  const double e = std::fma(s, kPow[decimals], -p);
  const double f = std::floor(p);
  const double frac = p - f;  // exact
  unsigned long long n = (unsigned long long)f;
  bool up;
  if (frac < 0.25) up = false;
  else if (frac > 0.75) up = true;
  else {
    const double c = (frac - 0.5) + e;  // frac - 0.5 is exact; the sign of the sum is exact
    up = c > 0.0 || (c == 0.0 && (n & 1ull));
  }
  n += up ? 1ull : 0ull;
  if (std::signbit(x)) *o++ = '-';
  o = put_int(o, (long long)(n / kPowI[decimals]));
  if (decimals > 0) {
    unsigned long long r = n % kPowI[decimals];
    *o++ = '.';
    for (int i = decimals - 1; i >= 0; i--) { o[i] = (char)('0' + r % 10); r /= 10; }
    o += decimals;
  }
  return o;
}
}  // namespace

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
  std::string row;
  format_site(row, chrom, hdr, r, persons, pr);
  write_rows(row.data(), row.size(), 1);
}

void VcfWriter::write_rows(const char *text, size_t bytes, long n_rows) {
  if (n_rows <= 0) return;
  if (!header_done_) header(opt_.denovo);
  fwrite(text, 1, bytes, fh_);
  fflush(fh_);
  rows_ += n_rows;
}

void VcfWriter::format_site(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                            const pm_person_site *persons, const pm_person_result *pr) const {
  if (opt_.denovo) format_denovo(out, chrom, hdr, r, persons, pr);
  else format_normal(out, chrom, hdr, r, persons, pr);
}

static inline int depth_of(const pm_person_site &p) { return p.depth[0] | (p.depth[1] << 8) | (p.depth[2] << 16); }

// CHROM .. INFO's common prefix "NS=..;PS=..;DP=..;MQ=.."
static void append_site_prefix(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                               const char *alt, size_t alt_len) {
  out += chrom; out.push_back('\t');
  append_int(out, (long long)hdr.pos + 1);
  out += "\t.\t"; out.push_back(kBases[hdr.ref_base]); out.push_back('\t');
  out.append(alt, alt_len); out.push_back('\t');
  append_int(out, (long long)int(r.poly_qual + 0.5));
  out += "\t.\tNS="; append_int(out, r.num_samp);
  out += ";PS="; append_fixed(out, r.perc_samp * 100, 1);
  out += ";DP="; append_int(out, r.total_depth);
  out += ";MQ="; append_fixed(out, r.avg_map_qual, 1);
}

void VcfWriter::format_normal(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                              const pm_person_site *persons, const pm_person_result *pr) const {
  const bool single_nuclear = ped_.families.size() == 1 && ped_.families[0].nuclear();
  const bool mono = (r.flags & PM_FLAG_MONO) != 0;
  const int a1 = r.allele1, a2 = r.allele2, ref = hdr.ref_base;
  const int np = ped_.n_person();
  out.reserve(out.size() + 160 + (size_t)np * 28);
  char alt[4];
  size_t alt_len;
  if (ref == a1) { alt[0] = kBases[mono ? a1 : a2]; alt_len = 1; }
  else { alt[0] = kBases[a1]; alt[1] = ','; alt[2] = kBases[a2]; alt_len = 3; }
  append_site_prefix(out, chrom, hdr, r, alt, alt_len);
  if (!single_nuclear) {
    out += ";AF="; append_fixed(out, r.freq, 4);
    if (hdr.chr_class == PM_CHR_AUTO) { out += ";AB="; append_fixed(out, r.ab, 3); }
  }
  if (mono) { out += ";BA="; out.push_back(kBases[a2]); }
  out += opt_.gl_off ? "\tGT:GQ:DP:DS" : "\tGT:GQ:DP:DS:PL";
  static const char *lab[5] = {"0/0", "0/1", "1/1", "1/2", "2/2"};
  static const char *lab_hap[5] = {"0", "ERROR", "1", "ERROR2", "2"};
  const int g11 = genotype_index(a1, a1), g12 = genotype_index(a1, a2), g22 = genotype_index(a2, a2);
  // GetBestGenoLabel_vcfv4 (NucFam.cpp:1587-1608): haploid labels on Y / MT and for males on X; "." for females on Y
  // (NucFam.cpp:626, 644, 663, 792; FLSeq.cpp:181-188)
  const bool hap_all = hdr.chr_class == PM_CHR_Y || hdr.chr_class == PM_CHR_MT;
  const bool chr_x = hdr.chr_class == PM_CHR_X, chr_y = hdr.chr_class == PM_CHR_Y;
  const std::vector<int> &cols = ped_.columns();
  // sample columns through a raw cursor: at most kPerPerson bytes each unless the dosage needs the snprintf fallback
  const size_t kPerPerson = 64;
  const size_t base = out.size();
  out.resize(base + (size_t)np * kPerPerson + kMaxFixed + 8);
  char *o = &out[base];
  const bool gl = !opt_.gl_off;
  for (int i = 0; i < np; i++) {
    if ((size_t)(o - &out[base]) + kPerPerson + kMaxFixed > out.size() - base) {  // a fallback-sized dosage ate the slack
      const size_t used = (size_t)(o - &out[0]);
      out.resize(out.size() + (size_t)(np - i) * kPerPerson + kMaxFixed);
      o = &out[used];
    }
    const int best = pr[i].best;
    const int label_idx = (ref == a1) ? best : best + 2;
    const int sex = (chr_x || chr_y) ? ped_.persons[cols[(size_t)i]].sex : 0;
    const char *l = (chr_y && sex == 2) ? "." : (hap_all || (chr_x && sex == 1)) ? lab_hap[label_idx] : lab[label_idx];
    *o++ = '\t';
    while (*l) *o++ = *l++;
    *o++ = ':'; o = put_u8(o, (unsigned)pr[i].gq & 0xff);
    *o++ = ':'; o = put_int(o, depth_of(persons[i]));
    *o++ = ':'; o = put_fixed(o, pr[i].dosage, 2);
    if (gl) {
      *o++ = ':'; o = put_u8(o, persons[i].lk[g11]);
      *o++ = ','; o = put_u8(o, persons[i].lk[g12]);
      *o++ = ','; o = put_u8(o, persons[i].lk[g22]);
    }
  }
  *o++ = '\n';
  out.resize((size_t)(o - &out[0]));
}

void VcfWriter::format_denovo(std::string &out, const std::string &chrom, const pm_site_hdr &hdr, const pm_site_result &r,
                              const pm_person_site *persons, const pm_person_result *pr) const {
  const bool single_nuclear = ped_.families.size() == 1 && ped_.families[0].nuclear();
  const bool mono = (r.flags & PM_FLAG_MONO) != 0;
  const int a1 = r.allele1, ref = hdr.ref_base;
  const int a2_label = r.allele2;           // alleles the genotype labels were made with
  const int a2 = mono ? a1 : r.allele2;     // denovo_mono: allele2 = allele1 (NucFam.cpp:1870)
  out.reserve(out.size() + 160 + (size_t)ped_.n_person() * 52);
  char alt[4];
  size_t alt_len;
  if (ref == a1) { alt[0] = kBases[a2]; alt_len = 1; }
  else { alt[0] = kBases[a1]; alt[1] = ','; alt[2] = kBases[a2]; alt_len = 3; }
  append_site_prefix(out, chrom, hdr, r, alt, alt_len);
  if (!single_nuclear) { out += ";AF="; append_fixed(out, r.freq, 4); }
  out += ";DQ="; append_fixed(out, r.denovo_lr, 3);
  out += opt_.gl_off ? "\tGT:GQ:DP" : "\tGT:GQ:DP:PL";
  static const char *lab[5] = {"0/0", "0/1", "1/1", "1/2", "2/2"};
  static const char *lab_hap[5] = {"0", "ERROR", "1", "ERROR2", "2"};
  const size_t base = out.size();
  out.resize(base + (size_t)ped_.n_person() * 72 + 8);  // "\tA/A:255:16777215:" + ten u8 and nine commas < 72 bytes
  char *o = &out[base];
  const bool gl = !opt_.gl_off;
  int col = 0;
  for (const Family &f : ped_.families) {
    const bool letters = (int)f.path.size() != f.founders;  // nuclear / extended families print base letters
    for (size_t j = 0; j < f.path.size(); j++, col++) {
      const pm_person_result &p = pr[col];
      const char *gt;
      // CalcPostProb_SinglePerson -> vcfv4 label; the member `sex` it consults on X is never set under --denovo
      if (!letters) {
        const int li = (ref == a1) ? p.best : p.best + 2;
        const bool yfemale = hdr.chr_class == PM_CHR_Y && ped_.persons[f.path[j]].sex == 2;
        gt = yfemale ? "." : (hdr.chr_class == PM_CHR_Y || hdr.chr_class == PM_CHR_MT) ? lab_hap[li] : lab[li];
      }
      else if (p.ten_state) gt = kGenoLabel[p.best];
      else {
        int idx = p.best == 0 ? genotype_index(a1, a1) : p.best == 1 ? genotype_index(a1, a2_label) : genotype_index(a2_label, a2_label);
        gt = kGenoLabel[idx];
      }
      *o++ = '\t';
      while (*gt) *o++ = *gt++;
      *o++ = ':'; o = put_u8(o, (unsigned)p.gq & 0xff);
      *o++ = ':'; o = put_int(o, depth_of(persons[col]));
      if (gl) {
        *o++ = ':';
        for (int g = 0; g < 9; g++) { o = put_u8(o, persons[col].lk[g]); *o++ = ','; }
        o = put_u8(o, persons[col].lk[9]);
      }
    }
  }
  *o++ = '\n';
  out.resize((size_t)(o - &out[0]));
}

}  // namespace pmh
