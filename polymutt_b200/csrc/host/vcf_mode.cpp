#include "vcf_mode.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "vcf_writer.h"

namespace pmh {

namespace {

int fatal(const std::string &msg) {
  printf("\nFATAL ERROR - \n%s\n\n", msg.c_str());
  return 1;
}

struct LineReader {  // plain or gzip, like base/IO.h's LineReader
  gzFile f = nullptr;
  std::vector<char> buf = std::vector<char>(1 << 20);
  bool open(const std::string &path) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 18); return f != nullptr; }
  bool next(std::string *line) {
    line->clear();
    for (;;) {
      if (!gzgets(f, buf.data(), (int)buf.size())) return !line->empty();
      size_t n = strlen(buf.data());
      line->append(buf.data(), n);
      if (n && (*line)[line->size() - 1] == '\n') {
        line->pop_back();
        if (!line->empty() && (*line)[line->size() - 1] == '\r') line->pop_back();
        return true;
      }
    }
  }
  ~LineReader() { if (f) gzclose(f); }
};

void split(const std::string &s, char sep, std::vector<std::string> *out) {
  out->clear();
  size_t b = 0;
  for (;;) {
    size_t e = s.find(sep, b);
    if (e == std::string::npos) { out->push_back(s.substr(b)); return; }
    out->push_back(s.substr(b, e - b));
    b = e + 1;
  }
}

// VCFRecord::getFormatIndex (libVcf/VCFRecord.h:283-309): prefix match at the start of each FORMAT field
int format_index(const std::string &format, const char *key) {
  size_t b = 0, e = format.size();
  int idx = 0;
  const size_t klen = strlen(key);
  while (b < e) {
    if (format.compare(b, klen, key) == 0) return idx;
    idx++;
    size_t c = format.find(':', b);
    if (c == std::string::npos) return -1;
    b = c + 1;
  }
  return -1;
}

int allele2int(const std::string &a) {  // FLSeq_VCF.cpp:65-72
  if (a == "A" || a == "a") return 1;
  if (a == "C" || a == "c") return 2;
  if (a == "G" || a == "g") return 3;
  if (a == "T" || a == "t") return 4;
  return 0;
}

struct Pending {  // one input record waiting for its output
  std::vector<std::string> col;      // the nine fixed columns
  std::vector<std::string> sample;   // included samples' whole fields, VCF order
  bool computed = false;             // false: printed with the state left by the previous computed record
  size_t row = 0;                    // row in the engine batch
};

}  // namespace

int run_vcf_mode(const Options &opt, const Pedigree &ped, const Engine &engine) {
  if (!engine.call_vcf) return fatal(std::string("engine '") + engine.name + "' has no VCF-input entry point");
  LineReader in;
  if (!in.open(opt.vcf_in)) return fatal("Cannot open VCF file " + opt.vcf_in);
  std::string line;
  std::vector<std::string> names;
  while (in.next(&line)) {
    if (line.rfind("##", 0) == 0) continue;
    if (line.rfind("#", 0) == 0) {
      std::vector<std::string> t;
      split(line, '\t', &t);
      if (t.size() <= 9) return fatal("not enough people in the VCF (VCF does not contain genotype and individuals?)");
      names.assign(t.begin() + 9, t.end());
      break;
    }
    return fatal("VCF header line (#CHROM ...) not found");
  }
  if (names.empty()) return fatal("VCF header line (#CHROM ...) not found");

  // MapPID2Traverse (FLSeq_VCF.cpp:38-56): pid -> column of the pedigree; a pid used in two families maps to the later one
  std::map<std::string, int> pid2col;
  {
    int c = 0;
    for (int idx : ped.columns()) pid2col[ped.persons[idx].pid] = c++;
  }
  std::vector<int> vcf2col(names.size(), -1);
  std::vector<std::string> included;
  int n_in_both = 0;
  for (size_t i = 0; i < names.size(); i++) {
    auto it = pid2col.find(names[i]);
    if (it == pid2col.end()) { printf("Sample ID \"%s\" not included in the analysis!\n", names[i].c_str()); continue; }
    vcf2col[i] = it->second;
    included.push_back(names[i]);
    n_in_both++;
  }

  FILE *out = fopen(opt.vcf_out.c_str(), "w");
  if (!out) return fatal("Open outpuf VCF file " + opt.vcf_out + " failed!");
  // meta data, PedVCF.cpp:82-102
  fprintf(out, "##fileformat=VCFv4.1\n##Polymutt=%s\n", opt.cmd.c_str());
  fprintf(out, "##Note=VCF file modified by polymutt. Updated fileds include: QUAL, GT and GQ, AF and AC. NOTE: modification was applied only to biallelic variants\n"
               "##FILTER=<ID=LOWDP,Description=\"Low Depth filter when the average depth per sample is lessn than 1\">\n"
               "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total Read Depth\">\n"
               "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Alternative Allele Frequency\">\n"
               "##INFO=<ID=AC,Number=1,Type=Integer,Description=\"Alternative Allele Count\">\n"
               "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
               "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">\n"
               "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read Depth\">\n"
               "##FORMAT=<ID=PL,Number=3,Type=Integer,Description=\"Phred-scaled Genotype Likelihoods\">\n"
               "##FORMAT=<ID=GL,Number=3,Type=Float,Description=\"Log10 Genotype Likelihoods\">\n");
  fprintf(out, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT");
  for (auto &n : included) fprintf(out, "\t%s", n.c_str());
  fprintf(out, "\n");

  pm_params par;
  opt.to_params(&par);
  par.vcf_input = 1;
  double lut[256];
  for (int i = 0; i < 256; i++) lut[i] = pow(10, -double(i) / 10.0);  // PL2LK_table, FLSeq_VCF.cpp:21-22
  void *ctx = engine.create(ped.view(), &par, lut, opt.device);
  if (!ctx) { fclose(out); return fatal(std::string("engine '") + engine.name + "': " + engine.last_error()); }

  const int np = ped.n_person();
  const size_t batch = opt.batch_sites > 0 ? (size_t)opt.batch_sites : (size_t)8192;
  std::vector<pm_site_hdr> hdr(batch);
  std::vector<pm_person_site> recs(batch * (size_t)np);
  std::vector<double> mono(batch);
  std::vector<pm_site_result> res(batch);
  std::vector<pm_person_result> pres(batch * (size_t)np);
  std::vector<Pending> pending;
  size_t n_rows = 0;

  // state that survives from record to record in the reference object (stale output for records without data)
  double last_qual = 0.0, last_min = 0.0;
  std::vector<int> last_best((size_t)np, 0), last_gq((size_t)np, 0);
  std::vector<char> last_labeled((size_t)np, 0);  // bestGenoLabel still "" until the first computed record
  int DP_index = -1, GL_idx = -1, PL_idx = -1;
  bool announced = false;

  auto flush = [&]() -> int {
    if (n_rows) {
      int rc = engine.call_vcf(ctx, hdr.data(), recs.data(), mono.data(), n_rows, res.data(), pres.data());
      if (rc != PM_OK) return rc;
    }
    std::vector<std::string> fd;
    for (const Pending &p : pending) {
      if (p.computed) {
        const pm_site_result &r = res[p.row];
        last_qual = r.poly_qual; last_min = r.freq;
        for (int c = 0; c < np; c++) {
          const pm_person_result &q = pres[p.row * (size_t)np + c];
          last_best[c] = q.best; last_gq[c] = q.gq; last_labeled[c] = 1;
        }
      }
      // FamilyLikelihoodSeq_VCF::OutputVCF, FLSeq_VCF.cpp:437-521
      int AC = 0, totalDepth = 0;
      bool missing = false;
      for (size_t i = 0, k = 0; i < names.size(); i++) {
        if (vcf2col[i] < 0) continue;
        const std::string &s = p.sample[k++];
        AC += last_best[vcf2col[i]];
        int dp = 0;
        if (DP_index > 0) {
          split(s, ':', &fd);
          missing = (size_t)DP_index >= fd.size() || fd[DP_index].empty();
          dp = missing ? 0 : atoi(fd[DP_index].c_str());
        }
        if (missing) continue;
        totalDepth += dp;
      }
      fprintf(out, "%s\t%d\t%s\t%s\t%s\t%.2f\t%s\tAF=%.2f;AC=%d;DP=%d\t%s", p.col[0].c_str(), atoi(p.col[1].c_str()), p.col[2].c_str(),
              p.col[3].c_str(), p.col[4].c_str(), last_qual, p.col[6].c_str(), 1 - last_min, AC, totalDepth,
              PL_idx > 0 ? "GT:GQ:DP:PL" : "GT:GQ:DP:GL");
      static const char *lab[3] = {"0/0", "0/1", "1/1"};
      for (size_t i = 0, k = 0; i < names.size(); i++) {
        if (vcf2col[i] < 0) continue;
        const std::string &s = p.sample[k++];
        const int c = vcf2col[i];
        split(s, ':', &fd);
        const char *label = last_labeled[c] ? lab[last_best[c]] : "";
        fprintf(out, "\t%s:%d:", last_gq[c] > 0 ? label : "./.", last_gq[c]);
        const char *dps = ".";
        if (DP_index > 0) {
          missing = (size_t)DP_index >= fd.size() || fd[DP_index].empty();
          dps = missing ? "" : fd[DP_index].c_str();
        }
        fprintf(out, "%s:", missing ? "." : dps);
        const int li = PL_idx > 0 ? PL_idx : GL_idx;
        missing = li < 0 || (size_t)li >= fd.size() || fd[li].empty();
        fprintf(out, "%s", missing ? "." : fd[li].c_str());
      }
      fprintf(out, "\n");
      fflush(out);
    }
    pending.clear();
    n_rows = 0;
    return PM_OK;
  };

  std::vector<std::string> t, fd, gl;
  int rc = PM_OK;
  while (in.next(&line)) {
    if (line.empty()) continue;
    split(line, '\t', &t);
    if (t.size() < 9 + names.size()) { engine.destroy(ctx); fclose(out); return fatal("VCF header have MORE people than VCF content!"); }
    if (!announced) { printf("Total samples in both VCF and PED files: %d\n\n", n_in_both); announced = true; }
    const std::string &refStr = t[3], &altStr = t[4];
    // FillPenetrance, FLSeq_VCF.cpp:267-383
    if (refStr == altStr) continue;                         // monomorphic: no output
    if (altStr.find(',') != std::string::npos) continue;    // not bi-allelic: no output
    const bool indel = refStr.size() > 1 || altStr.size() > 1;
    const int ref = indel ? 1 : allele2int(refStr), alt = indel ? 2 : allele2int(altStr);
    if (ref == 0 || alt == 0) {
      // the reference indexes its genotype table with Allele2Int() == 0 here (undefined behaviour); skipped instead
      printf("WARNING - REF/ALT %s/%s at %s:%s is not A, C, G or T; record skipped\n", refStr.c_str(), altStr.c_str(), t[0].c_str(), t[1].c_str());
      continue;
    }
    if (DP_index < 0) DP_index = format_index(t[8], "DP");
    if (GL_idx < 0 && PL_idx < 0) {
      GL_idx = format_index(t[8], "GL");
      PL_idx = format_index(t[8], "PL");
      if (GL_idx < 0 && PL_idx < 0) {
        fprintf(stderr, "NO GL or PL field was found. Please check the vcf file at chr:%s and position:%d", t[0].c_str(), atoi(t[1].c_str()));
        engine.destroy(ctx); fclose(out);
        return 1;
      }
      if (n_in_both == 0) { engine.destroy(ctx); fclose(out); return fatal("NO individual IDs match in the ped and vcf file!"); }
    }
    Pending p;
    p.col.assign(t.begin(), t.begin() + 9);
    for (size_t i = 0; i < names.size(); i++) if (vcf2col[i] >= 0) p.sample.push_back(t[9 + i]);
    pm_person_site *row = &recs[n_rows * (size_t)np];
    memset(row, 0, sizeof(pm_person_site) * (size_t)np);
    std::vector<double> loglk_rr((size_t)np, 0.0);
    const int g0 = genotype_index(ref, ref), g1 = genotype_index(ref, alt), g2 = genotype_index(alt, alt);
    int withdata = 0;
    for (size_t i = 0; i < names.size(); i++) {
      const int c = vcf2col[i];
      if (c < 0) continue;
      split(t[9 + i], ':', &fd);
      const int li = GL_idx > 0 ? GL_idx : PL_idx;
      const bool missing = li < 0 || (size_t)li >= fd.size() || fd[li].empty();
      if (missing) break;  // the reference returns from FillPenetrance here: later samples keep likelihood 1
      split(fd[li], ',', &gl);
      if (gl.size() != 3) {
        engine.destroy(ctx); fclose(out);
        return fatal("GL or PL filed does not have 3 values separated by commas at: " + t[0] + " " + t[1] + "!");
      }
      const double G[3] = {atof(gl[0].c_str()), atof(gl[1].c_str()), atof(gl[2].c_str())};
      if (G[0] != 0.0 || G[1] != 0.0 || G[2] != 0.0) withdata++;
      const int gi[3] = {g0, g1, g2};
      for (int k = 0; k < 3; k++) {
        const double ll = PL_idx > 0 ? (G[k] > 255 ? -255 / 10.0 : -G[k] / 10.0) : (-10 * G[k] > 255 ? -255 / 10.0 : G[k]);
        if (k == 0) loglk_rr[(size_t)c] = ll;
        int pl = int(PL_idx > 0 ? G[k] : -10 * G[k]);
        if (pl < 0) {
          engine.destroy(ctx); fclose(out);
          return fatal("Phred-scaled likelihood " + std::to_string(pl) + " can not be negative");
        }
        if (pl > 255) pl = 255;
        row[c].lk[gi[k]] = (uint8_t)pl;
      }
    }
    if (withdata == 0) {  // PedVCF.cpp:122: printed with whatever the previous record left behind
      pending.push_back(std::move(p));
      if (pending.size() >= 4 * batch) { if ((rc = flush()) != PM_OK) break; }
      continue;
    }
    // MonomorphismLogLikelihood, FLSeq_VCF.cpp:74-83: pedigree order
    double m = 0.0;
    for (int c = 0; c < np; c++) m += loglk_rr[(size_t)c];
    mono[n_rows] = m;
    hdr[n_rows].pos = (uint32_t)atoi(t[1].c_str());
    hdr[n_rows].ref_base = (uint8_t)ref;
    hdr[n_rows].chr_class = t[0] == opt.chrX ? PM_CHR_X : t[0] == opt.chrY ? PM_CHR_Y : t[0] == opt.chrMT ? PM_CHR_MT : PM_CHR_AUTO;
    hdr[n_rows].reserved = (uint16_t)(alt | (indel ? 0x100 : 0));
    p.computed = true;
    p.row = n_rows++;
    pending.push_back(std::move(p));
    if (n_rows == batch) { if ((rc = flush()) != PM_OK) break; }
  }
  if (rc == PM_OK) rc = flush();
  std::string err = rc == PM_OK ? std::string() : std::string("engine '") + engine.name + "': " + engine.last_error();
  engine.destroy(ctx);
  fclose(out);
  if (rc != PM_OK) return fatal(err);
  return 0;
}

}  // namespace pmh
