// pm-tools: fixture plumbing around the host front end (no GPU needed).
//
//   pm-tools pack   -p PED -d DAT -g GIF -o OUT.pmpk     pedigree view + every merged site, as the
//                                                        engine would receive them
//   pm-tools unpack IN.pmpk OUTDIR [--gzip]              the inverse: one GLF v3 file per VCF column,
//                                                        plus OUTDIR/ped, dat, gif text files
//
// .pmpk layout (little endian):
//   "PMPK" u32 version(1) i32 n_fam i32 n_person i32 n_steps
//   i32 fam_size[n_fam] i32 fam_founders[n_fam] i32 fam_generations[n_fam]
//   u8 sex[n_person] (padded to 4) i32 father[n_person] i32 mother[n_person] i32 glf_index[n_person]
//   i32 peel_first[n_fam+1]  pm_peel_step peel[n_steps]
//   u32 text_len, text: "famid pid fatid motid sex glf_index\n" per column, then "#label <chrom label>\n"
//   i32 max_position  u64 n_sites  pm_site_hdr hdr[n_sites]  pm_person_site rec[n_sites*n_person]
#include <sys/stat.h>
#include <time.h>
#include <cmath>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "glf.h"
#include "glf_ingest.h"
#include "params.h"
#include "vcf_writer.h"
#include "pedigree.h"

using namespace pmh;

static int usage() {
  fprintf(stderr, "usage: pm-tools pack -p PED -d DAT -g GIF -o OUT.pmpk | pm-tools unpack IN.pmpk OUTDIR [--gzip]\n");
  return 2;
}

template <typename T>
static void put(FILE *f, const T *p, size_t n) { fwrite(p, sizeof(T), n, f); }

static int do_pack(int argc, char **argv) {
  std::string ped_path, dat_path, gif_path, out_path;
  int batched = -1;  // --batched N: use the multi-threaded batch reader with N threads (0 = all cores)
  for (int i = 2; i + 1 < argc; i += 2) {
    std::string k = argv[i];
    if (k == "--batched") batched = atoi(argv[i + 1]);
    if (k == "--portable") GlfBatchReader::use_portable_convert(atoi(argv[i + 1]) != 0);  // --portable 1: no SSSE3 in the record conversion
    if (k == "-p") ped_path = argv[i + 1];
    else if (k == "-d") dat_path = argv[i + 1];
    else if (k == "-g") gif_path = argv[i + 1];
    else if (k == "-o") out_path = argv[i + 1];
  }
  if (ped_path.empty() || dat_path.empty() || out_path.empty()) return usage();
  Pedigree ped;
  try { ped.load(dat_path, ped_path); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  const pm_pedigree *v = ped.view();
  FILE *f = fopen(out_path.c_str(), "wb");
  if (!f) { perror("open output"); return 1; }
  const int32_t n_steps = v->peel_first[v->n_fam];
  uint32_t version = 1;
  fwrite("PMPK", 1, 4, f);
  put(f, &version, 1); put(f, &v->n_fam, 1); put(f, &v->n_person, 1); put(f, &n_steps, 1);
  put(f, v->fam_size, v->n_fam); put(f, v->fam_founders, v->n_fam); put(f, v->fam_generations, v->n_fam);
  std::vector<uint8_t> sex(((size_t)v->n_person + 3) / 4 * 4, 0);
  memcpy(sex.data(), v->sex, (size_t)v->n_person);
  put(f, sex.data(), sex.size());
  put(f, v->father, v->n_person); put(f, v->mother, v->n_person);
  std::vector<int32_t> glf_index;
  std::ostringstream text;
  for (int idx : ped.columns()) {
    const Person &p = ped.persons[idx];
    glf_index.push_back(p.glf_index);
    text << p.famid << ' ' << p.pid << ' ' << p.fatid << ' ' << p.motid << ' ' << p.sex << ' ' << p.glf_index << '\n';
  }
  put(f, glf_index.data(), glf_index.size());
  put(f, v->peel_first, (size_t)v->n_fam + 1);
  if (n_steps) put(f, v->peel, (size_t)n_steps);

  std::vector<pm_site_hdr> hdrs;
  std::vector<pm_person_site> recs;
  std::string label;
  int32_t max_position = 0;
  if (!gif_path.empty()) {
    std::map<std::string, std::string> glf_map;
    std::ifstream g(gif_path);
    std::string line;
    while (std::getline(g, line)) {
      std::istringstream in(line);
      std::string a, b;
      if (in >> a >> b) glf_map[a] = b;
    }
    std::vector<std::string> paths;
    for (int gi : glf_index) paths.push_back(gi == 0 || !glf_map.count(std::to_string(gi)) ? std::string() : glf_map[std::to_string(gi)]);
    std::string err;
    if (batched >= 0) {
      GlfBatchReader glf;
      if (!glf.open(paths, batched, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
      if (glf.next_section()) {
        label = glf.label();
        max_position = glf.max_position();
        const size_t B = 1000;  // deliberately not a power of two: batch boundaries fall inside runs
        std::vector<pm_site_hdr> hb(B);
        std::vector<pm_person_site> rb(B * (size_t)v->n_person);
        size_t n;
        while ((n = glf.next_batch(hb.data(), rb.data(), B)) > 0) {
          hdrs.insert(hdrs.end(), hb.begin(), hb.begin() + (long)n);
          recs.insert(recs.end(), rb.begin(), rb.begin() + (long)(n * (size_t)v->n_person));
        }
      }
    } else {
      GlfSet glf;
      if (!glf.open(paths, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
      if (glf.next_section()) {  // fixtures hold one chromosome
        label = glf.label();
        max_position = glf.max_position();
        pm_site_hdr h;
        std::vector<pm_person_site> one((size_t)v->n_person);
        while (glf.next_site(&h, one.data())) { hdrs.push_back(h); recs.insert(recs.end(), one.begin(), one.end()); }
      }
    }
  }
  text << "#label " << label << '\n';
  std::string t = text.str();
  uint32_t tl = (uint32_t)t.size();
  put(f, &tl, 1); fwrite(t.data(), 1, t.size(), f);
  put(f, &max_position, 1);
  uint64_t ns = hdrs.size();
  put(f, &ns, 1);
  put(f, hdrs.data(), hdrs.size());
  put(f, recs.data(), recs.size());
  fclose(f);
  fprintf(stderr, "packed %llu sites x %d persons -> %s\n", (unsigned long long)ns, v->n_person, out_path.c_str());
  return 0;
}

template <typename T>
static bool get(FILE *f, T *p, size_t n) { return fread(p, sizeof(T), n, f) == n; }

static int do_unpack(int argc, char **argv) {
  if (argc < 4) return usage();
  const std::string in = argv[2], outdir = argv[3];
  const bool gzip = argc > 4 && std::string(argv[4]) == "--gzip";
  FILE *f = fopen(in.c_str(), "rb");
  if (!f) { perror("open input"); return 1; }
  char magic[4];
  uint32_t version;
  int32_t n_fam, n_person, n_steps;
  if (!get(f, magic, 4) || memcmp(magic, "PMPK", 4) || !get(f, &version, 1) || !get(f, &n_fam, 1) || !get(f, &n_person, 1) || !get(f, &n_steps, 1)) {
    fprintf(stderr, "not a .pmpk file\n");
    return 1;
  }
  std::vector<int32_t> skip((size_t)n_fam * 3);
  get(f, skip.data(), skip.size());
  std::vector<uint8_t> sex(((size_t)n_person + 3) / 4 * 4);
  get(f, sex.data(), sex.size());
  std::vector<int32_t> fm((size_t)n_person * 2), glf_index((size_t)n_person), pf((size_t)n_fam + 1);
  get(f, fm.data(), fm.size()); get(f, glf_index.data(), glf_index.size()); get(f, pf.data(), pf.size());
  std::vector<pm_peel_step> steps((size_t)n_steps);
  if (n_steps) get(f, steps.data(), steps.size());
  uint32_t tl;
  get(f, &tl, 1);
  std::string text(tl, '\0');
  get(f, &text[0], tl);
  int32_t max_position;
  uint64_t ns;
  get(f, &max_position, 1); get(f, &ns, 1);
  std::vector<pm_site_hdr> hdrs(ns);
  std::vector<pm_person_site> recs(ns * (size_t)n_person);
  if (!get(f, hdrs.data(), hdrs.size()) || !get(f, recs.data(), recs.size())) { fprintf(stderr, "truncated .pmpk\n"); return 1; }
  fclose(f);
  mkdir(outdir.c_str(), 0755);
  std::string label = "1";
  std::ofstream ped(outdir + "/ped"), dat(outdir + "/dat"), gif(outdir + "/gif");
  dat << "T\tGLF_Index\n";
  std::istringstream tin(text);
  std::string line;
  int col = 0;
  std::set<int> seen;
  while (std::getline(tin, line)) {
    if (line.rfind("#label ", 0) == 0) { label = line.substr(7); continue; }
    std::istringstream in(line);
    std::string famid, pid, fat, mot; int sx, gi;
    in >> famid >> pid >> fat >> mot >> sx >> gi;
    // every distinct GLF_Index gets its own file, keyed by that index (so other pedigrees can
    // refer to the same streams the way the original GLF index file did)
    int key = gi;
    ped << famid << '\t' << pid << '\t' << fat << '\t' << mot << '\t' << sx << '\t' << key << '\n';
    if (key && !seen.count(key)) gif << key << ' ' << outdir << "/col" << key << ".glf\n";
    seen.insert(key);
    col++;
  }
  std::set<int> written;
  for (int c = 0; c < n_person; c++) {
    const int key = glf_index[(size_t)c];
    if (key == 0 || written.count(key)) continue;
    written.insert(key);
    GlfWriter w;
    std::string path = outdir + "/col" + std::to_string(key) + ".glf";
    if (!w.create(path, gzip)) { perror(path.c_str()); return 1; }
    w.begin_section(label, max_position);
    for (uint64_t s = 0; s < ns; s++) {
      const pm_person_site &r = recs[s * (size_t)n_person + (size_t)c];
      uint32_t depth = r.depth[0] | (r.depth[1] << 8) | (r.depth[2] << 16);
      // a person without a record at this position is all zeros with depth 0: leave the position out
      bool empty = depth == 0 && r.map_quality == 0;
      for (int g = 0; g < 10 && empty; g++) empty = r.lk[g] == 0;
      if (empty) continue;
      w.write_entry((int)hdrs[s].pos, hdrs[s].ref_base, depth, r.map_quality, r.lk);
    }
    w.end_section();
    w.close();
  }
  fprintf(stderr, "unpacked %llu sites x %d persons -> %s\n", (unsigned long long)ns, n_person, outdir.c_str());
  return 0;
}

// pm-tools ingest-bench -p PED -d DAT -g GIF [--batched N] [--batch B]: reads everything, reports sites/s of the merge alone
static int do_ingest_bench(int argc, char **argv) {
  std::string ped_path, dat_path, gif_path;
  int batched = -1;
  size_t B = 4096;
  for (int i = 2; i + 1 < argc; i += 2) {
    std::string k = argv[i];
    if (k == "-p") ped_path = argv[i + 1];
    else if (k == "-d") dat_path = argv[i + 1];
    else if (k == "-g") gif_path = argv[i + 1];
    else if (k == "--batched") batched = atoi(argv[i + 1]);
    else if (k == "--batch") B = (size_t)atol(argv[i + 1]);
  }
  Pedigree ped;
  try { ped.load(dat_path, ped_path); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  std::map<std::string, std::string> glf_map;
  std::ifstream g(gif_path);
  std::string line;
  while (std::getline(g, line)) { std::istringstream in(line); std::string a, b; if (in >> a >> b) glf_map[a] = b; }
  std::vector<std::string> paths;
  for (int idx : ped.columns()) {
    int gi = ped.persons[idx].glf_index;
    paths.push_back(gi == 0 || !glf_map.count(std::to_string(gi)) ? std::string() : glf_map[std::to_string(gi)]);
  }
  const size_t np = paths.size();
  std::vector<pm_site_hdr> hb(B);
  std::vector<pm_person_site> rb(B * np);
  std::string err;
  size_t total = 0;
  unsigned long long check = 0;
  struct timespec t0, t1, t2;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (batched >= 0) {
    GlfBatchReader glf;
    if (!glf.open(paths, batched, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    while (glf.next_section()) { size_t n; while ((n = glf.next_batch(hb.data(), rb.data(), B)) > 0) { total += n; check += hb[n - 1].pos + rb[(n - 1) * np].lk[0]; } }
  } else {
    GlfSet glf;
    if (!glf.open(paths, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    while (glf.next_section()) { while (glf.next_site(&hb[0], rb.data())) { total++; check += hb[0].pos + rb[0].lk[0]; } }
  }
  clock_gettime(CLOCK_MONOTONIC, &t2);
  auto sec = [](const timespec &a, const timespec &b) { return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec); };
  printf("{\"sites\": %zu, \"persons\": %zu, \"open_s\": %.3f, \"merge_s\": %.3f, \"sites_per_s\": %.0f, \"packed_MB_per_s\": %.0f, \"check\": %llu}\n",
         total, np, sec(t0, t1), sec(t1, t2), (double)total / sec(t1, t2), (double)total * (double)np * 16 / 1e6 / sec(t1, t2), check);
  return 0;
}

// pm-tools fmt-selftest [N]: append_fixed / append_int against snprintf on random values, exact ties and edge cases
static int do_fmt_selftest(int argc, char **argv) {
  const long N = argc > 2 ? atol(argv[2]) : 1000000;
  unsigned long long st = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  long bad = 0, done = 0;
  auto check = [&](double x, int d) {
    char buf[400];
    snprintf(buf, sizeof buf, "%.*f", d, x);
    std::string s;
    append_fixed(s, x, d);
    done++;
    if (s != buf) { if (bad < 10) fprintf(stderr, "mismatch %.17g d=%d: %s vs %s\n", x, d, s.c_str(), buf); bad++; }
  };
  const double edge[] = {0.0, -0.0, 0.5, 1.5, 2.5, 0.125, 0.375, 0.005, 0.015, 0.025, 0.045, 1e-300, 4.9e-324, 0.99999, 0.999999999,
                         99.95, 99.949999999999, 100.0, 1e8, 9.99999999e8, 1e9, 1e15, 1e300, -0.0001, -3.14159, 2.0, 1.0 / 3.0};
  for (double x : edge) for (int d = 0; d <= 6; d++) { check(x, d); check(-x, d); }
  check(NAN, 2); check(INFINITY, 3); check(-INFINITY, 1);
  for (long i = 0; i < N; i++) {
    const unsigned long long r = rnd();
    const int d = (int)(r % 5);
    double x;
    switch ((r >> 8) % 5) {
      case 0: x = (double)(rnd() >> 11) / 9007199254740992.0 * 2.0; break;               // dosage-like
      case 1: x = (double)(rnd() % 2000001) / 1000.0 / 2.0; break;                         // many exact decimal ties
      case 2: x = ((double)(rnd() % 200001) + 0.5) / std::pow(10.0, d); break;             // nearest doubles to ties
      case 3: x = std::ldexp((double)(rnd() >> 11), (int)(rnd() % 80) - 90); break;        // wide exponent range
      default: x = -(double)(rnd() >> 11) / 9007199254740992.0 * 300.0; break;             // negative log ratios
    }
    check(x, d);
    check(std::nextafter(x, 1e300), d);
    check(std::nextafter(x, -1e300), d);
  }
  long ibad = 0;
  for (long i = 0; i < 100000; i++) {
    long long v = (long long)rnd() >> (int)(rnd() % 63);
    if (i & 1) v = -v;
    char buf[32];
    snprintf(buf, sizeof buf, "%lld", v);
    std::string s;
    append_int(s, v);
    if (s != buf) ibad++;
  }
  printf("{\"checked\": %ld, \"fixed_mismatches\": %ld, \"int_mismatches\": %ld}\n", done, bad, ibad);
  return bad || ibad ? 1 : 0;
}

// pm-tools fmt-bench -p PED -d DAT [--denovo] [--rows N]: VCF row formatting speed on made-up results
static int do_fmt_bench(int argc, char **argv) {
  std::string ped_path, dat_path;
  long rows = 2000;
  Options opt;
  for (int i = 2; i < argc; i++) {
    std::string k = argv[i];
    if (k == "--denovo") opt.denovo = true;
    else if (i + 1 < argc) {
      if (k == "-p") ped_path = argv[++i];
      else if (k == "-d") dat_path = argv[++i];
      else if (k == "--rows") rows = atol(argv[++i]);
    }
  }
  Pedigree ped;
  try { ped.load(dat_path, ped_path); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  const int np = ped.n_person();
  VcfWriter w(nullptr, opt, ped);
  std::vector<pm_person_site> ps((size_t)np);
  std::vector<pm_person_result> pr((size_t)np);
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  for (int i = 0; i < np; i++) {
    for (int g = 0; g < 10; g++) ps[(size_t)i].lk[g] = (uint8_t)(rnd() % 256);
    ps[(size_t)i].depth[0] = (uint8_t)(rnd() % 40);
    pr[(size_t)i].best = (uint8_t)(rnd() % 3); pr[(size_t)i].gq = (uint8_t)(rnd() % 101); pr[(size_t)i].dosage = (double)(rnd() % 2001) / 1000.0;
  }
  pm_site_hdr h; memset(&h, 0, sizeof h); h.pos = 123456; h.ref_base = 1;
  pm_site_result r; memset(&r, 0, sizeof r);
  r.allele1 = 1; r.allele2 = 3; r.poly_qual = 57.3; r.num_samp = np; r.perc_samp = 0.98; r.total_depth = 45000; r.avg_map_qual = 59.2; r.freq = 0.9312; r.ab = 0.48;
  std::string out;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  size_t bytes = 0;
  for (long i = 0; i < rows; i++) { out.clear(); w.format_site(out, "1", h, r, ps.data(), pr.data()); bytes += out.size(); }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  printf("{\"rows\": %ld, \"persons\": %d, \"rows_per_s\": %.0f, \"MB_per_s\": %.0f, \"bytes_per_row\": %zu}\n", rows, np, (double)rows / sec, (double)bytes / 1e6 / sec, bytes / (size_t)rows);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) return usage();
  if (!strcmp(argv[1], "fmt-bench")) return do_fmt_bench(argc, argv);
  if (!strcmp(argv[1], "fmt-selftest")) return do_fmt_selftest(argc, argv);
  if (!strcmp(argv[1], "ingest-bench")) return do_ingest_bench(argc, argv);
  if (!strcmp(argv[1], "pack")) return do_pack(argc, argv);
  if (!strcmp(argv[1], "unpack")) return do_unpack(argc, argv);
  return usage();
}
