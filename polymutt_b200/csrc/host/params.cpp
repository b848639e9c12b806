#include "params.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pmh {

namespace {
enum Kind { K_BOOL, K_INT, K_DOUBLE, K_STRING };
struct LongOpt { const char *name; Kind kind; void *ptr; };

bool is_int(const char *s) {  // core/Parameters.cpp CheckInteger
  if (*s == '+' || *s == '-') s++;
  if (!*s) return false;
  for (; *s; s++) if (!isdigit((unsigned char)*s)) return false;
  return true;
}
bool is_double(const char *s) {
  char *end = nullptr;
  strtod(s, &end);
  return end != s && *end == 0;
}
bool ieq_prefix(const char *stem, const char *name) {
  for (; *stem; stem++, name++)
    if (tolower((unsigned char)*stem) != tolower((unsigned char)*name)) return false;
  return true;
}
}  // namespace

bool Options::parse(int argc, char **argv, std::string *err) {
  LongOpt table[] = {
      {"in_vcf", K_STRING, &vcf_in},        {"theta", K_DOUBLE, &theta},
      {"indel_theta", K_DOUBLE, &theta_indel}, {"poly_tstv", K_DOUBLE, &tstv_ratio},
      {"chrX", K_STRING, &chrX},            {"chrY", K_STRING, &chrY},
      {"MT", K_STRING, &chrMT},             {"denovo", K_BOOL, &denovo},
      {"rate_denovo", K_DOUBLE, &denovo_mut_rate}, {"tstv_denovo", K_DOUBLE, &denovo_tstv_ratio},
      {"minLLR_denovo", K_DOUBLE, &denovo_lr},     {"prec", K_DOUBLE, &precision},
      {"nthreads", K_INT, &num_threads},    {"chr2process", K_STRING, &chrs2process},
      {"minMapQuality", K_INT, &min_map_quality},  {"minDepth", K_INT, &min_total_depth},
      {"maxDepth", K_INT, &max_total_depth},       {"minPercSampleWithData", K_DOUBLE, &min_ps},
      {"out_vcf", K_STRING, &vcf_out},      {"pos", K_STRING, &pos_file},
      {"all_sites", K_BOOL, &out_all_sites},       {"gl_off", K_BOOL, &gl_off},
      {"quick_call", K_BOOL, &quick_call},
      // ours
      {"device", K_INT, &device},           {"batch_sites", K_INT, &batch_sites},
      {"gpus", K_INT, &gpus},               {"ingest_threads", K_INT, &ingest_threads},
  };
  const int n_long = (int)(sizeof table / sizeof table[0]);
  cmd.clear();
  for (int a = 0; a < argc; a++) { cmd += argv[a]; cmd += " "; }

  auto set_value = [&](LongOpt &o, const char *v) {
    switch (o.kind) {
      case K_INT: *(int *)o.ptr = atoi(v); break;
      case K_DOUBLE: *(double *)o.ptr = atof(v); break;
      case K_STRING: *(std::string *)o.ptr = v; break;
      case K_BOOL: break;
    }
  };
  for (int i = 1; i < argc; i++) {
    const char *arg = argv[i];
    bool ok = false;
    if (arg[0] == '-' && arg[1] == '-') {
      std::string name(arg + 2), inline_value;
      bool has_inline = false;
      size_t colon = name.find(':');
      if (colon != std::string::npos) { inline_value = name.substr(colon + 1); name.resize(colon); has_inline = true; }
      int hit = -1, hits = 0;
      for (int j = 0; j < n_long; j++) {
        if (strcasecmp(name.c_str(), table[j].name) == 0) { hit = j; hits = 1; break; }
        if (!name.empty() && ieq_prefix(name.c_str(), table[j].name)) { hit = j; hits++; }
      }
      if (hits == 1) {
        LongOpt &o = table[hit];
        if (o.kind == K_BOOL) { *(bool *)o.ptr = !*(bool *)o.ptr; ok = true; }
        else if (has_inline) { set_value(o, inline_value.c_str()); ok = true; }
        else if (i + 1 < argc && ((o.kind == K_INT && is_int(argv[i + 1])) || (o.kind == K_DOUBLE && is_double(argv[i + 1])) || o.kind == K_STRING)) {
          set_value(o, argv[++i]); ok = true;
        }
      } else if (hits > 1) {
        warnings.push_back(std::string("Ambiguous Option: Command line parameter ") + arg + " matches several options\n");
        continue;
      }
    } else if (arg[0] == '-' && arg[1]) {
      char ch = (char)tolower((unsigned char)arg[1]);
      std::string *sp = ch == 'p' ? &ped_file : ch == 'd' ? &dat_file : ch == 'g' ? &glf_index_file : nullptr;
      if (sp || ch == 'c') {
        const char *v = nullptr;
        if (arg[2] == 0 && i + 1 < argc && argv[i + 1][0] != '-') v = argv[++i];
        else v = arg + 2;
        if (sp) *sp = v; else posterior = atof(v);
        ok = true;
      }
    }
    if (!ok) {
      char buf[512];
      snprintf(buf, sizeof buf, "Command line parameter %s (#%d) ignored\n", arg, i);
      warnings.push_back(buf);
    }
  }
  // main.cpp:139-153
  if (vcf_in == vcf_out) { *err = "Input and output VCF files are the same!"; return false; }
  if (ped_file.empty()) { *err = "pedFile not provided for input!"; return false; }
  if (glf_index_file.empty() && vcf_in.empty()) { *err = "glfListFile or input VCF file not provided for input!"; return false; }
  if (vcf_out.empty()) { *err = "vcfOutFile not provided for output!"; return false; }
  if (!pos_file.empty()) { force_call = true; quick_call = false; out_all_sites = false; }
  if (out_all_sites) quick_call = false;
  if (denovo && denovo_lr < 0) { *err = "denovo_min_LLR can only be greater than 0 !"; return false; }
  return true;
}

void Options::to_params(pm_params *p) const {
  memset(p, 0, sizeof *p);
  p->theta = theta; p->theta_indel = theta_indel; p->poly_tstv = tstv_ratio;
  p->posterior_cutoff = posterior; p->precision = precision;
  p->denovo_mut_rate = denovo_mut_rate; p->denovo_tstv = denovo_tstv_ratio; p->denovo_min_llr = denovo_lr;
  p->min_ps = min_ps; p->min_map_quality = min_map_quality;
  p->min_total_depth = min_total_depth; p->max_total_depth = max_total_depth;
  p->denovo = denovo; p->force_call = force_call; p->out_all_sites = out_all_sites; p->quick_call = quick_call;
}

void Options::print_status() const {
  printf("\nThe following parameters are in effect:\n");
  printf("                  pedfile : %s (-pname)\n", ped_file.c_str());
  printf("                  datfile : %s (-dname)\n", dat_file.c_str());
  printf("             glfIndexFile : %s (-gname)\n", glf_index_file.c_str());
  printf("         posterior cutoff : %.3f (-c99.999)\n", posterior);
  printf("\nAdditional Options\n");
  printf("  Alternative input file : --in_vcf [%s]\n", vcf_in.c_str());
  printf("    Scaled mutation rate : --theta [%.1e], --indel_theta [%.1e]\n", theta, theta_indel);
  printf("  Prior of ts/tv ratio : --poly_tstv [%.2f]\n", tstv_ratio);
  printf("   Non-autosome labels : --chrX [%s], --chrY [%s], --MT [%s]\n", chrX.c_str(), chrY.c_str(), chrMT.c_str());
  printf("      de novo mutation : --denovo%s, --rate_denovo [%.1e], --tstv_denovo [%.2f], --minLLR_denovo [%.2f]\n",
         denovo ? " [ON]" : "", denovo_mut_rate, denovo_tstv_ratio, denovo_lr);
  printf("Optimization precision : --prec [%.1e]\n", precision);
  printf("    Multiple threading : --nthreads [%d]\n", num_threads);
  printf("Chromosomes to process : --chr2process [%s]\n", chrs2process.c_str());
  printf("               Filters : --minMapQuality [%d], --minDepth [%d], --maxDepth [%d], --minPercSampleWithData [%.2f]\n",
         min_map_quality, min_total_depth, max_total_depth, min_ps);
  printf("                Output : --out_vcf [%s], --pos [%s], --all_sites%s, --gl_off%s, --quick_call%s\n\n",
         vcf_out.c_str(), pos_file.c_str(), out_all_sites ? " [ON]" : "", gl_off ? " [ON]" : "", quick_call ? " [ON]" : "");
  if (!warnings.empty()) {
    printf("WARNING - Problems encountered parsing command line:\n\n");
    for (auto &w : warnings) printf("%s", w.c_str());
    printf("\n");
  }
}

}  // namespace pmh
