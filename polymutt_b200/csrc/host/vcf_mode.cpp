#include "vcf_mode.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "vcf_writer.h"

namespace pmh {

namespace {

int fatal(const std::string &msg) {
  printf("\nFATAL ERROR - \n%s\n\n", msg.c_str());
  return 1;
}

struct LineReader {  // plain or gzip, like base/IO.h's LineReader; used for the header lines only
  gzFile f = nullptr;
  std::vector<char> buf = std::vector<char>(1 << 20);
  std::string rest;  // always empty: gzgets never reads past the line it returns
  bool open(const std::string &path) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
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

// ---- allocation-free tokenising of a chunk of lines (libVcf LINE_MODE rules, see vcf_mode.h) ----
struct Tok {
  const char *p = nullptr;
  uint32_t n = 0;
  bool eq(const Tok &o) const { return n == o.n && memcmp(p, o.p, n) == 0; }
  bool has(char c) const { return n && memchr(p, c, n) != nullptr; }
  std::string str() const { return std::string(p, n); }
};

// k-th `sep`-separated field of t; false if there are fewer than k+1 fields
inline bool nth_field(const Tok &t, char sep, int k, Tok *out) {
  const char *b = t.p, *end = t.p + t.n;
  for (int i = 0; i < k; i++) {
    const char *c = (const char *)memchr(b, sep, (size_t)(end - b));
    if (!c) return false;
    b = c + 1;
  }
  const char *c = (const char *)memchr(b, sep, (size_t)(end - b));
  out->p = b;
  out->n = (uint32_t)((c ? c : end) - b);
  return true;
}

// fields ia and ib (':'-separated, -1 = none) of a sample token in one pass; has_x = the token has that many fields
inline void two_fields(const Tok &t, int ia, int ib, Tok *fa, bool *has_a, Tok *fb, bool *has_b) {
  *has_a = *has_b = false;
  const char *end = t.p + t.n, *fs = t.p;
  int k = 0;
  for (const char *q = t.p;; q++) {
    if (q == end || *q == ':') {
      if (k == ia) { fa->p = fs; fa->n = (uint32_t)(q - fs); *has_a = true; }
      if (k == ib) { fb->p = fs; fb->n = (uint32_t)(q - fs); *has_b = true; }
      if (q == end) break;
      k++; fs = q + 1;
      if (k > ia && k > ib) break;
    }
  }
}

// atof() of a token that is not NUL-terminated
inline double tok_atof(const Tok &t) {
  if (t.n == 0) return 0.0;
  if (t.n <= 15) {  // plain digits (every PL): exact
    unsigned long long v = 0;
    uint32_t i = 0;
    for (; i < t.n && t.p[i] >= '0' && t.p[i] <= '9'; i++) v = v * 10 + (unsigned)(t.p[i] - '0');
    if (i == t.n) return (double)v;
  }
  char buf[64];
  if (t.n < sizeof buf) { memcpy(buf, t.p, t.n); buf[t.n] = 0; return atof(buf); }
  return atof(t.str().c_str());
}
inline int tok_atoi(const Tok &t) {
  char buf[32];
  const uint32_t n = t.n < sizeof buf - 1 ? t.n : (uint32_t)sizeof buf - 1;
  memcpy(buf, t.p, n);
  buf[n] = 0;
  return atoi(buf);
}

int format_index(const Tok &format, const char *key) { return format_index(format.str(), key); }
int allele2int(const Tok &a) { return a.n == 1 ? allele2int(std::string(1, a.p[0])) : 0; }

enum LineKind : uint8_t { L_SKIP = 0, L_WARN, L_NODATA, L_COMPUTED, L_ERROR };

struct LineRec {
  Tok line;
  Tok col[9];
  size_t samp0 = 0;      // first of this line's names.size() sample tokens in the chunk's token array
  uint8_t kind = L_SKIP;
  uint8_t ref = 0, alt = 0, indel = 0;
  int dp_here = -1;      // format_index(FORMAT, "DP") of this line
  int dp_index = -1;     // the DP index in force when this line is printed (set in line order)
  double mono = 0.0;
  long row = -1;         // engine row (computed lines)
  long src = -1;         // engine row whose results this line prints; -1 = state carried over from earlier chunks
  std::string err;       // L_ERROR / L_WARN text
};

// Reads the (plain or gzip) file in large blocks and hands out chunks of whole lines.  The next block is read (and
// inflated) by a helper thread while the current chunk is tokenised, computed and formatted.
struct ChunkReader {
  gzFile f = nullptr;
  std::vector<char> buf;
  size_t have = 0;       // bytes in buf
  size_t start = 0;      // first unconsumed byte
  bool eof = false;
  static constexpr size_t kBlock = (size_t)24 << 20;  // a chunk's worth of text
  std::vector<char> ahead;      // the block being read ahead
  size_t ahead_n = 0;
  bool ahead_eof = false;
  std::thread ahead_thread;
  bool open(const std::string &path) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
  ~ChunkReader() { if (ahead_thread.joinable()) ahead_thread.join(); if (f) gzclose(f); }
  void read_ahead() {
    ahead.resize(kBlock);
    ahead_thread = std::thread([this]() {
      size_t n = 0;
      while (n < kBlock) {
        const int got = gzread(f, ahead.data() + n, (unsigned)(kBlock - n));
        if (got <= 0) { ahead_eof = true; break; }
        n += (size_t)got;
      }
      ahead_n = n;
    });
  }
  // Next line (without its terminator); false at end of input.  Pointers stay valid until the next refill().
  bool next_line(Tok *out, bool allow_refill) {
    for (;;) {
      const char *b = buf.data() + start;
      const char *nl = have > start ? (const char *)memchr(b, '\n', have - start) : nullptr;
      if (nl) {
        size_t n = (size_t)(nl - b);
        start += n + 1;
        if (n && b[n - 1] == '\r') n--;
        out->p = b; out->n = (uint32_t)n;
        return true;
      }
      if (eof) {
        if (have == start) return false;
        size_t n = have - start;
        start = have;
        if (n && b[n - 1] == '\r') n--;
        out->p = b; out->n = (uint32_t)n;
        return true;
      }
      if (!allow_refill) return false;
      refill();
    }
  }
  // Drops consumed bytes and appends the block the helper has read; invalidates every pointer handed out before.
  void refill() {
    if (start > 0) { memmove(buf.data(), buf.data() + start, have - start); have -= start; start = 0; }
    if (!eof) {
      if (!ahead_thread.joinable()) read_ahead();  // the first block
      ahead_thread.join();
      if (buf.size() < have + ahead_n + 1) buf.resize(have + ahead_n + 1);
      memcpy(buf.data() + have, ahead.data(), ahead_n);
      have += ahead_n;
      ahead_n = 0;
      if (ahead_eof) eof = true; else read_ahead();
    }
    if (buf.size() < have + 1) buf.resize(have + 1);
    buf[have] = 0;
  }
};

// The buffers that cross the engine boundary: page-locked when the engine offers an allocator (the CUDA engine's copies
// are then asynchronous and at link speed), plain memory otherwise.  The few std::vector members the loop uses.
template <typename T>
struct HostVec {
  const Engine &e;
  T *p = nullptr;
  size_t n = 0, cap = 0;
  explicit HostVec(const Engine &eng) : e(eng) {}
  HostVec(const HostVec &) = delete;
  ~HostVec() { release(p); }
  void release(T *q) { if (!q) return; if (e.host_alloc && e.host_free) e.host_free(q); else free(q); }
  void resize(size_t m) {
    if (m > cap) {
      const size_t want = std::max(m, cap + cap / 2);
      T *q = (T *)((e.host_alloc && e.host_free) ? e.host_alloc(want * sizeof(T)) : malloc(want * sizeof(T)));
      if (!q) throw std::bad_alloc();
      if (n) memcpy(q, p, n * sizeof(T));
      release(p);
      p = q; cap = want;
    }
    n = m;
  }
  size_t size() const { return n; }
  T *data() { return p; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

template <typename F>
void parallel_for(size_t n, int threads, F fn) {
  if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) fn(i, 0); return; }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  const int nt = (int)std::min<size_t>((size_t)threads, n);
  for (int t = 0; t < nt; t++)
    pool.emplace_back([&, t]() {
      for (;;) {
        const size_t i0 = next.fetch_add(16);
        if (i0 >= n) break;
        for (size_t i = i0; i < std::min(n, i0 + 16); i++) fn(i, t);
      }
    });
  for (auto &th : pool) th.join();
}

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
  // One engine context per GPU (--gpus N, ours): the records of a chunk are independent, so each chunk's rows are cut
  // into N contiguous ranges, one per context, each on its own host thread; results land in place, in record order.
  const int n_gpu = opt.gpus > 0 ? opt.gpus : 1;
  std::vector<void *> ctxs((size_t)n_gpu, nullptr);
  auto destroy_all = [&]() { for (void *c : ctxs) if (c) engine.destroy(c); ctxs.clear(); };
  {
    std::vector<std::string> ctx_err((size_t)n_gpu);
    std::vector<std::thread> makers;
    auto make = [&](int g) {
      ctxs[(size_t)g] = engine.create(ped.view(), &par, lut, opt.device + g);
      if (!ctxs[(size_t)g]) ctx_err[(size_t)g] = engine.last_error();  // the message is thread-local: keep it
    };
    for (int g = 1; g < n_gpu; g++) makers.emplace_back(make, g);
    make(0);
    for (auto &t : makers) t.join();
    for (int g = 0; g < n_gpu; g++)
      if (!ctxs[(size_t)g]) { destroy_all(); fclose(out); return fatal(std::string("engine '") + engine.name + "': " + ctx_err[(size_t)g]); }
  }
  std::string engine_err;
  // runs call(ctx, first row, number of rows) over the chunk's rows, one contiguous range per context
  auto sharded = [&](size_t n_rows, const std::function<int(void *, size_t, size_t)> &call) -> int {
    const size_t n_use = std::min<size_t>((size_t)n_gpu, n_rows);
    if (n_use <= 1) {
      const int r = call(ctxs[0], 0, n_rows);
      if (r != PM_OK) engine_err = engine.last_error();
      return r;
    }
    std::vector<int> rcs(n_use, PM_OK);
    std::vector<std::string> errs(n_use);
    std::vector<std::thread> workers;
    auto part = [&](size_t g) {
      const size_t lo = n_rows * g / n_use, hi = n_rows * (g + 1) / n_use;
      rcs[g] = call(ctxs[g], lo, hi - lo);
      if (rcs[g] != PM_OK) errs[g] = engine.last_error();
    };
    for (size_t g = 1; g < n_use; g++) workers.emplace_back(part, g);
    part(0);
    for (auto &t : workers) t.join();
    for (size_t g = 0; g < n_use; g++) if (rcs[g] != PM_OK) { engine_err = errs[g]; return rcs[g]; }
    return PM_OK;
  };

  const int np = ped.n_person();
  const size_t n_names = names.size();
  int threads = opt.ingest_threads > 0 ? opt.ingest_threads : (int)std::thread::hardware_concurrency();
  threads = std::max(1, std::min(threads, 32));
  const size_t max_lines = opt.batch_sites > 0 ? (size_t)opt.batch_sites : (size_t)8192;

  // state that survives from record to record in the reference object (stale output for records without data)
  double last_qual = 0.0, last_min = 0.0;
  std::vector<int> last_best((size_t)np, 0), last_gq((size_t)np, 0);
  bool last_labeled = false;  // bestGenoLabel is still "" until the first computed record
  int last_cls = PM_CHR_AUTO; // chromosome class of the record those labels were made on (haploid / "." labels)
  std::vector<int> col_sex((size_t)np, 0);
  {
    int c = 0;
    for (int idx : ped.columns()) col_sex[(size_t)c++] = ped.persons[idx].sex;
  }
  int DP_index = -1, GL_idx = -1, PL_idx = -1;
  bool announced = false;

  std::vector<LineRec> lines;
  std::vector<Tok> toks;
  HostVec<pm_site_hdr> hdr(engine);
  // per (line, pedigree column): the three PL bytes of the record's genotypes (a1a1, a1a2, a2a2); engines without the
  // pl3 entry point (the CPU oracle behind the same front end) get them widened to 16-byte records below
  HostVec<uint8_t> pl3(engine);
  std::vector<pm_person_site> recs;
  HostVec<double> mono(engine);
  HostVec<pm_site_result> res(engine);
  std::vector<pm_person_result> pres;   // per-sample results, or (engines with call_vcf_calls) ...
  HostVec<uint16_t> calls(engine);       // ... best | gq << 8 per sample
  const bool compact = engine.call_vcf_calls != nullptr || engine.call_vcf_pl != nullptr;
  std::vector<std::string> text, text_w;  // rows being formatted / rows being written
  std::thread writer;
  std::vector<std::vector<double>> scratch((size_t)threads, std::vector<double>((size_t)np));
  struct SampF { Tok dp, lk; bool has_dp, has_lk; };  // the DP and PL / GL fields of one sample token
  std::vector<std::vector<SampF>> fscratch((size_t)threads, std::vector<SampF>(n_names));
  static const char *lab[3] = {"0/0", "0/1", "1/1"};

  // FillPenetrance for one line (FLSeq_VCF.cpp:267-383); everything it writes belongs to the line
  auto parse_line = [&](size_t li, int tid) {
    LineRec &L = lines[li];
    if (L.kind == L_ERROR) return;
    const Tok &refStr = L.col[3], &altStr = L.col[4];
    if (refStr.eq(altStr)) { L.kind = L_SKIP; return; }   // monomorphic: no output
    if (altStr.has(',')) { L.kind = L_SKIP; return; }      // not bi-allelic: no output
    const bool indel = refStr.n > 1 || altStr.n > 1;
    const int ref = indel ? 1 : allele2int(refStr), alt = indel ? 2 : allele2int(altStr);
    if (ref == 0 || alt == 0) {
      // the reference indexes its genotype table with Allele2Int() == 0 here (undefined behaviour); skipped instead
      L.kind = L_WARN;
      L.err = "WARNING - REF/ALT " + refStr.str() + "/" + altStr.str() + " at " + L.col[0].str() + ":" + L.col[1].str() + " is not A, C, G or T; record skipped";
      return;
    }
    L.ref = (uint8_t)ref; L.alt = (uint8_t)alt; L.indel = indel;
    L.dp_here = format_index(L.col[8], "DP");
    uint8_t *row = &pl3[li * (size_t)np * 3];
    memset(row, 0, (size_t)np * 3);
    std::vector<double> &loglk_rr = scratch[(size_t)tid];
    std::fill(loglk_rr.begin(), loglk_rr.end(), 0.0);
    const int fi = GL_idx > 0 ? GL_idx : PL_idx;
    int withdata = 0;
    for (size_t i = 0; i < n_names; i++) {
      const int c = vcf2col[i];
      if (c < 0) continue;
      double G[3];
      // fast path: the field is three runs of decimal digits "a,b,c" (every PL a caller writes): one pass over the
      // sample's bytes, no memchr calls; anything else (GL floats, empty or missing values, a fourth value) goes the
      // general way below with the same result
      bool fast = false;
      if (fi >= 0) {
        const Tok &st = toks[L.samp0 + i];
        const char *b = st.p, *end = st.p + st.n;
        int k = 0;
        while (k < fi && b < end) if (*b++ == ':') k++;
        if (k == fi && b < end && *b != ':') {
          unsigned v[3] = {0, 0, 0};
          int idx = 0, digits = 0;
          bool ok = true;
          for (const char *q = b; q < end && *q != ':'; q++) {
            const unsigned ch = (unsigned char)*q;
            if (ch - '0' <= 9u) { v[idx] = v[idx] * 10 + (ch - '0'); if (++digits > 8) { ok = false; break; } }
            else if (ch == ',' && digits > 0 && idx < 2) { idx++; digits = 0; }
            else { ok = false; break; }
          }
          if (ok && idx == 2 && digits > 0) { G[0] = (double)v[0]; G[1] = (double)v[1]; G[2] = (double)v[2]; fast = true; }
        }
      }
      if (!fast) {
        Tok field, g[3], extra;
        const bool missing = fi < 0 || !nth_field(toks[L.samp0 + i], ':', fi, &field) || field.n == 0;
        if (missing) break;  // the reference returns from FillPenetrance here: later samples keep likelihood 1
        if (!nth_field(field, ',', 0, &g[0]) || !nth_field(field, ',', 1, &g[1]) || !nth_field(field, ',', 2, &g[2]) || nth_field(field, ',', 3, &extra)) {
          L.kind = L_ERROR;
          L.err = "GL or PL filed does not have 3 values separated by commas at: " + L.col[0].str() + " " + L.col[1].str() + "!";
          return;
        }
        G[0] = tok_atof(g[0]); G[1] = tok_atof(g[1]); G[2] = tok_atof(g[2]);
      }
      if (G[0] != 0.0 || G[1] != 0.0 || G[2] != 0.0) withdata++;
      for (int k = 0; k < 3; k++) {
        const double ll = PL_idx > 0 ? (G[k] > 255 ? -255 / 10.0 : -G[k] / 10.0) : (-10 * G[k] > 255 ? -255 / 10.0 : G[k]);
        if (k == 0) loglk_rr[(size_t)c] = ll;
        int pl = int(PL_idx > 0 ? G[k] : -10 * G[k]);
        if (pl < 0) { L.kind = L_ERROR; L.err = "Phred-scaled likelihood " + std::to_string(pl) + " can not be negative"; return; }
        if (pl > 255) pl = 255;
        row[3 * c + k] = (uint8_t)pl;
      }
    }
    if (withdata == 0) { L.kind = L_NODATA; return; }  // PedVCF.cpp:122: printed with whatever the previous record left behind
    // MonomorphismLogLikelihood, FLSeq_VCF.cpp:74-83: pedigree order
    double m = 0.0;
    for (int c = 0; c < np; c++) m += loglk_rr[(size_t)c];
    L.mono = m;
    L.kind = L_COMPUTED;
  };

  // FamilyLikelihoodSeq_VCF::OutputVCF (FLSeq_VCF.cpp:437-521) for one line, into text[li]
  auto format_line = [&](size_t li, int tid) {
    const LineRec &L = lines[li];
    std::string &o = text[li];
    o.clear();
    if (L.kind != L_NODATA && L.kind != L_COMPUTED) return;
    const bool from_row = L.src >= 0;
    const double qual = from_row ? res[(size_t)L.src].poly_qual : last_qual;
    const double fmin = from_row ? res[(size_t)L.src].freq : last_min;
    const pm_person_result *pr = from_row && !compact ? &pres[(size_t)L.src * (size_t)np] : nullptr;
    const uint16_t *cl = from_row && compact ? &calls[(size_t)L.src * (size_t)np] : nullptr;
    const bool labeled = from_row || last_labeled;
    // GetBestGenoLabel_vcfv4 (NucFam.cpp:1587-1608) on the record the labels were made on: haploid labels on Y / MT and
    // for males on X, "." for females on Y (FLSeq_VCF.cpp:204, 221-227)
    const int lcls = from_row ? (int)hdr[(size_t)L.src].chr_class : last_cls;
    static const char *lab_hap[3] = {"0", "ERROR", "1"};
    auto best_of = [&](int c) { return pr ? (int)pr[c].best : (cl ? (int)(cl[c] & 0xff) : last_best[(size_t)c]); };
    auto gq_of = [&](int c) { return pr ? (int)pr[c].gq : (cl ? (int)(cl[c] >> 8) : last_gq[(size_t)c]); };
    const int dpi = L.dp_index;
    const int fi = PL_idx > 0 ? PL_idx : GL_idx;
    int AC = 0, totalDepth = 0;
    bool missing = false;
    // the DP and PL / GL fields of every sample, located once (one pass over the sample's bytes)
    std::vector<SampF> &sf = fscratch[(size_t)tid];
    for (size_t i = 0; i < n_names; i++) {
      if (vcf2col[i] < 0) continue;
      SampF &x = sf[i];
      two_fields(toks[L.samp0 + i], dpi > 0 ? dpi : -1, fi >= 0 ? fi : -1, &x.dp, &x.has_dp, &x.lk, &x.has_lk);
      AC += best_of(vcf2col[i]);
      int dp = 0;
      if (dpi > 0) {
        missing = !x.has_dp || x.dp.n == 0;
        dp = missing ? 0 : tok_atoi(x.dp);
      }
      if (missing) continue;
      totalDepth += dp;
    }
    o.reserve(L.line.n + 64);
    o.append(L.col[0].p, L.col[0].n); o.push_back('\t');
    append_int(o, tok_atoi(L.col[1])); o.push_back('\t');
    o.append(L.col[2].p, L.col[2].n); o.push_back('\t');
    o.append(L.col[3].p, L.col[3].n); o.push_back('\t');
    o.append(L.col[4].p, L.col[4].n); o.push_back('\t');
    append_fixed(o, qual, 2); o.push_back('\t');
    o.append(L.col[6].p, L.col[6].n);
    o += "\tAF="; append_fixed(o, 1 - fmin, 2);
    o += ";AC="; append_int(o, AC);
    o += ";DP="; append_int(o, totalDepth);
    o += PL_idx > 0 ? "\tGT:GQ:DP:PL" : "\tGT:GQ:DP:GL";
    // the sample columns through a raw pointer: label ':' GQ ':' DP ':' PL, at most 24 bytes beside the copied fields
    const size_t head = o.size();
    o.resize(head + L.line.n + 24 * n_names + 8);
    char *w = &o[head];
    for (size_t i = 0; i < n_names; i++) {
      if (vcf2col[i] < 0) continue;
      const int c = vcf2col[i];
      const SampF &x = sf[i];
      const int gq = gq_of(c);
      *w++ = '\t';
      const int sex = col_sex[(size_t)c];
      const char *label = !labeled ? "" : (lcls == PM_CHR_Y && sex == 2) ? "."
                          : (lcls == PM_CHR_Y || lcls == PM_CHR_MT || (lcls == PM_CHR_X && sex == 1)) ? lab_hap[best_of(c)] : lab[best_of(c)];
      const char *shown = (gq > 0 || (label[0] == '.' && label[1] == 0)) ? label : "./.";   // FLSeq_VCF.cpp:507
      while (*shown) *w++ = *shown++;
      *w++ = ':';
      {  // GQ (0..255 from the engine; any int from the stale state)
        char tmp[12];
        int n = 0;
        unsigned v = gq < 0 ? 0u - (unsigned)gq : (unsigned)gq;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        if (gq < 0) *w++ = '-';
        while (n) *w++ = tmp[--n];
      }
      *w++ = ':';
      Tok dps; dps.p = "."; dps.n = 1;
      if (dpi > 0) {
        missing = !x.has_dp || x.dp.n == 0;
        if (!missing) dps = x.dp;
      }
      if (missing) *w++ = '.'; else { memcpy(w, dps.p, dps.n); w += dps.n; }
      *w++ = ':';
      missing = fi < 0 || !x.has_lk || x.lk.n == 0;
      if (missing) *w++ = '.'; else { memcpy(w, x.lk.p, x.lk.n); w += x.lk.n; }
    }
    o.resize((size_t)(w - o.data()));
    o.push_back('\n');
  };

  ChunkReader rd;
  rd.f = in.f; in.f = nullptr;      // continue where the header scan stopped
  rd.buf.assign(in.rest.begin(), in.rest.end());
  rd.have = rd.buf.size();
  rd.buf.resize(rd.have + ((size_t)32 << 20));
  int rc = PM_OK;
  std::string fail;
  bool more = true;
  double tm[5] = {0, 0, 0, 0, 0};  // PM_TIMING: read, tokenise + parse, engine, format, write
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  while (more && rc == PM_OK && fail.empty()) {
    const auto t_begin = std::chrono::steady_clock::now();
    // ---- one chunk of whole lines, tokenised by tabs ----
    if (!rd.eof && rd.have - rd.start < ((size_t)4 << 20)) rd.refill();
    lines.clear();
    Tok line;
    while (lines.size() < max_lines && rd.next_line(&line, false)) {
      if (line.n == 0) continue;
      LineRec L;
      L.line = line;
      lines.push_back(std::move(L));
    }
    if (lines.empty()) {
      if (rd.eof && rd.have == rd.start) { more = false; break; }
      if (!rd.eof) rd.refill();  // a line longer than what was buffered: the buffer grows as needed
      continue;
    }
    const size_t nl = lines.size();
    const auto t_read = std::chrono::steady_clock::now();
    toks.assign(nl * n_names, Tok());
    parallel_for(nl, threads, [&](size_t li, int) {
      LineRec &L = lines[li];
      L.samp0 = li * n_names;
      const char *b = L.line.p, *end = L.line.p + L.line.n;
      size_t k = 0;
      while (k < 9 + n_names) {
        const char *c = (const char *)memchr(b, '\t', (size_t)(end - b));
        Tok t; t.p = b; t.n = (uint32_t)((c ? c : end) - b);
        if (k < 9) L.col[k] = t; else toks[L.samp0 + (k - 9)] = t;
        k++;
        if (!c) break;
        b = c + 1;
      }
      if (k < 9 + n_names) { L.kind = L_ERROR; L.err = "VCF header have MORE people than VCF content!"; }
    });
    // ---- the FORMAT indices are fixed by the first record that reaches them (FLSeq_VCF.cpp:340-350) ----
    if (GL_idx < 0 && PL_idx < 0) {
      for (size_t li = 0; li < nl; li++) {
        const LineRec &L = lines[li];
        if (L.kind == L_ERROR) break;
        if (L.col[3].eq(L.col[4]) || L.col[4].has(',')) continue;
        const bool indel = L.col[3].n > 1 || L.col[4].n > 1;
        if (!indel && (allele2int(L.col[3]) == 0 || allele2int(L.col[4]) == 0)) continue;
        GL_idx = format_index(L.col[8], "GL");
        PL_idx = format_index(L.col[8], "PL");
        if (GL_idx < 0 && PL_idx < 0) {
          fprintf(stderr, "NO GL or PL field was found. Please check the vcf file at chr:%s and position:%d", L.col[0].str().c_str(), tok_atoi(L.col[1]));
          if (writer.joinable()) writer.join();
          destroy_all(); fclose(out);
          return 1;
        }
        if (n_in_both == 0) { if (writer.joinable()) writer.join(); destroy_all(); fclose(out); return fatal("NO individual IDs match in the ped and vcf file!"); }
        break;
      }
    }
    if (!announced) { printf("Total samples in both VCF and PED files: %d\n\n", n_in_both); announced = true; }
    if (pl3.size() < nl * (size_t)np * 3) pl3.resize(nl * (size_t)np * 3);
    parallel_for(nl, threads, parse_line);
    const auto t_parse = std::chrono::steady_clock::now();
    // ---- in line order: warnings, the first error, engine rows, which row each line prints ----
    size_t n_rows = 0, n_use = nl;
    long src = -1;
    hdr.resize(nl); mono.resize(nl);
    for (size_t li = 0; li < nl; li++) {
      LineRec &L = lines[li];
      if (L.kind == L_WARN) { printf("%s\n", L.err.c_str()); continue; }
      if (L.kind == L_ERROR) { fail = L.err; n_use = li; break; }
      if (L.kind == L_SKIP) continue;
      if (DP_index < 0) DP_index = L.dp_here;
      L.dp_index = DP_index;
      if (L.kind == L_COMPUTED) {
        L.row = (long)n_rows;
        if (n_rows != li) memmove(&pl3[n_rows * (size_t)np * 3], &pl3[li * (size_t)np * 3], (size_t)np * 3);
        mono[n_rows] = L.mono;
        pm_site_hdr &h = hdr[n_rows];
        memset(&h, 0, sizeof h);
        h.pos = (uint32_t)tok_atoi(L.col[1]);
        h.ref_base = L.ref;
        const std::string chrom = L.col[0].str();
        h.chr_class = chrom == opt.chrX ? PM_CHR_X : chrom == opt.chrY ? PM_CHR_Y : chrom == opt.chrMT ? PM_CHR_MT : PM_CHR_AUTO;
        h.reserved = (uint16_t)(L.alt | (L.indel ? 0x100 : 0));
        src = (long)n_rows++;
      }
      L.src = src;
    }
    if (n_rows) {
      if (res.size() < n_rows) { res.resize(n_rows); if (compact) calls.resize(n_rows * (size_t)np); else pres.resize(n_rows * (size_t)np); }
      if (engine.call_vcf_pl) {
        rc = sharded(n_rows, [&](void *c, size_t lo, size_t n) {
          return engine.call_vcf_pl(c, hdr.data() + lo, pl3.data() + lo * (size_t)np * 3, mono.data() + lo, n, res.data() + lo, calls.data() + lo * (size_t)np);
        });
      } else {
        if (recs.size() < n_rows * (size_t)np) recs.resize(n_rows * (size_t)np);
        parallel_for(n_rows, threads, [&](size_t r, int) {
          const int a1 = hdr[r].ref_base, a2 = hdr[r].reserved & 0xff;
          const int gi[3] = {genotype_index(a1, a1), genotype_index(a1, a2), genotype_index(a2, a2)};
          pm_person_site *row = &recs[r * (size_t)np];
          memset(row, 0, sizeof(pm_person_site) * (size_t)np);
          for (int c = 0; c < np; c++)
            for (int k = 0; k < 3; k++) row[c].lk[gi[k]] = pl3[(r * (size_t)np + (size_t)c) * 3 + (size_t)k];
        });
        rc = sharded(n_rows, [&](void *c, size_t lo, size_t n) {
          const size_t at = lo * (size_t)np;
          return compact ? engine.call_vcf_calls(c, hdr.data() + lo, recs.data() + at, mono.data() + lo, n, res.data() + lo, calls.data() + at)
                         : engine.call_vcf(c, hdr.data() + lo, recs.data() + at, mono.data() + lo, n, res.data() + lo, pres.data() + at);
        });
      }
      if (rc != PM_OK) break;
    }
    const auto t_engine = std::chrono::steady_clock::now();
    if (text.size() < nl) text.resize(nl);
    parallel_for(n_use, threads, format_line);
    const auto t_format = std::chrono::steady_clock::now();
    // the rows of this chunk are written by a helper thread while the next chunk is read, parsed and computed
    if (writer.joinable()) writer.join();
    text.swap(text_w);
    writer = std::thread([&text_w, n_use, out]() {
      for (size_t li = 0; li < n_use; li++) if (!text_w[li].empty()) fwrite(text_w[li].data(), 1, text_w[li].size(), out);
      fflush(out);
    });
    const auto t_write = std::chrono::steady_clock::now();
    tm[0] += secs(t_begin, t_read); tm[1] += secs(t_read, t_parse); tm[2] += secs(t_parse, t_engine); tm[3] += secs(t_engine, t_format); tm[4] += secs(t_format, t_write);
    if (n_rows) {  // what the next chunk's leading no-data records print
      const size_t r = n_rows - 1;
      last_qual = res[r].poly_qual; last_min = res[r].freq; last_labeled = true; last_cls = hdr[r].chr_class;
      for (int c = 0; c < np; c++) {
        last_best[(size_t)c] = compact ? (int)(calls[r * (size_t)np + c] & 0xff) : pres[r * (size_t)np + c].best;
        last_gq[(size_t)c] = compact ? (int)(calls[r * (size_t)np + c] >> 8) : pres[r * (size_t)np + c].gq;
      }
    }
  }
  if (writer.joinable()) writer.join();
  if (getenv("PM_TIMING"))
    printf("[pm timing] vcf mode: read %.3f s, tokenise+parse %.3f s, engine %.3f s, format %.3f s, write %.3f s; %d threads\n", tm[0], tm[1], tm[2], tm[3], tm[4], threads);
  std::string err = rc == PM_OK ? std::string() : std::string("engine '") + engine.name + "': " + engine_err;
  destroy_all();
  fclose(out);
  if (rc != PM_OK) return fatal(err);
  if (!fail.empty()) return fatal(fail);
  return 0;
}

}  // namespace pmh
