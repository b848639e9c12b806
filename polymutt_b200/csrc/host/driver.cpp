#include "driver.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <future>
#include <memory>
#include <mutex>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "glf.h"
#include "glf_ingest.h"
#include "params.h"
#include "pedigree.h"
#include "vcf_mode.h"
#include "vcf_writer.h"

namespace pmh {

namespace {

int fatal(const std::string &msg) {  // core/Error.cpp:10-23
  printf("\nFATAL ERROR - \n%s\n\n", msg.c_str());
  return 1;
}

std::vector<std::string> split_ws(const std::string &line) {
  std::vector<std::string> out;
  std::istringstream in(line);
  std::string t;
  while (in >> t) out.push_back(t);
  return out;
}

// .gif: "<index> <glf path>" per line (src/main.cpp:15-37)
bool read_glf_index(const std::string &path, std::map<std::string, std::string> *m, std::string *err) {
  std::ifstream f(path);
  if (!f) { *err = path + " open failed"; return false; }
  std::string line;
  while (std::getline(f, line)) {
    auto tok = split_ws(line);
    if (tok.size() < 2) continue;
    (*m)[tok[0]] = tok[1];
  }
  return true;
}

struct Counters {  // src/main.cpp:264-282
  unsigned minTotalDepthFilter = 0, maxTotalDepthFilter = 0, minMapQualFilter = 0, minPSFilter = 0;
  int refBaseCounts[5] = {0, 0, 0, 0, 0};
  int homoRef = 0, transitions = 0, transversions = 0, tstvs1 = 0, tstvs2 = 0, tvs1tvs2 = 0, nocall = 0;
  int totalEntryCnt = 0;
};

// Counter updates of main.cpp:341-553 from one status word (see pm_call_glf_sites).
void count_site(Counters &c, const pm_site_hdr &h, uint16_t st, const Options &opt) {
  int code = st & 0xf, maxidx = ((st >> 4) & 0xf) - 1;
  bool nocall = (st >> 8) & 1;
  if (code == PM_SITE_BAD_REF) return;
  c.refBaseCounts[h.ref_base]++;
  switch (code) {
    case PM_SITE_MIN_DEPTH: c.minTotalDepthFilter++; return;
    case PM_SITE_MAX_DEPTH: c.maxTotalDepthFilter++; return;
    case PM_SITE_MIN_PS: c.minPSFilter++; return;
    case PM_SITE_MIN_MAPQ: c.minMapQualFilter++; return;
    case PM_SITE_QUICK_SKIP: return;
  }
  if (nocall) { c.nocall++; if (!opt.force_call && !opt.out_all_sites) return; }
  switch (maxidx) {
    case 0: c.homoRef++; break;
    case 1: c.transitions++; break;
    case 2: case 3: c.transversions++; break;
    case 4: c.tstvs1++; break;
    case 5: c.tstvs2++; break;
    case 6: c.tvs1tvs2++; break;
  }
}

}  // namespace

int run_cli(int argc, char **argv, const Engine &engine) {
  const double t_start = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  Options opt;
  std::string err;
  bool parsed = opt.parse(argc, argv, &err);
  opt.print_status();
  if (!parsed) return fatal(err);
  Pedigree ped;
  try {
    ped.load(opt.dat_file, opt.ped_file);
  } catch (const std::exception &e) {
    return fatal(e.what());
  }
  if (!opt.vcf_in.empty()) return run_vcf_mode(opt, ped, engine);  // main.cpp:238-246
  auto wall = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_ped_done = wall();

  std::map<std::string, std::string> glf_map;
  if (!read_glf_index(opt.glf_index_file, &glf_map, &err)) return fatal(err);
  FILE *vcf = fopen(opt.vcf_out.c_str(), "w");
  if (!vcf) return fatal("vcfOutFile can not be opened for output!");

  // PedigreeGLF::SetPedGLF, src/PedigreeGLF.cpp:117-163
  std::vector<std::string> paths;
  for (int idx : ped.columns()) {
    const Person &p = ped.persons[idx];
    if (p.glf_index == 0) { paths.emplace_back(); continue; }
    auto it = glf_map.find(std::to_string(p.glf_index));
    if (it == glf_map.end()) {
      printf("\nWARNING - \nNo entry found for the glf with the key [%d]\n\n", p.glf_index);
      paths.emplace_back();
      continue;
    }
    paths.push_back(it->second);
  }
  // The engine contexts (CUDA initialisation, module load: the longest single step of the set-up) are created on a
  // background thread while the GLF files are opened and their headers read.
  pm_params par;
  opt.to_params(&par);
  double lut[256];
  pm_fill_lut(lut);
  // One engine context per GPU (--gpus N, ours): batches of consecutive sites go to the GPUs round-robin, each on
  // its own host thread, and are consumed strictly in site order, so the VCF is the ordered concatenation of the
  // per-GPU shards.  No communication between GPUs.
  const int n_gpu = opt.gpus > 0 ? opt.gpus : 1;
  std::vector<void *> ctxs;
  auto destroy_all = [&]() { for (void *c : ctxs) engine.destroy(c); ctxs.clear(); };
  std::string ctx_error;
  double t_ctx = 0.0;
  std::future<bool> ctx_ready = std::async(std::launch::async, [&]() -> bool {
    const double t0 = wall();
    for (int g = 0; g < n_gpu; g++) {
      void *c = engine.create(ped.view(), &par, lut, opt.device + g);
      if (!c) { ctx_error = engine.last_error(); return false; }
      ctxs.push_back(c);
    }
    t_ctx = wall() - t0;
    return true;
  });
  GlfBatchReader glf;  // multi-threaded block decode + merge (glf_ingest.h); GlfSet in glf.h is the one-site-at-a-time form
  const bool glf_ok = glf.open(paths, opt.ingest_threads, &err);
  const double t_open_done = wall();
  const bool ctx_ok = ctx_ready.get();
  if (!glf_ok) { destroy_all(); return fatal(err); }
  if (!ctx_ok) { destroy_all(); return fatal(std::string("engine '") + engine.name + "': " + ctx_error); }

  std::set<std::string> positions;  // --pos: "chr:pos" (src/main.cpp:39-55)
  if (!opt.pos_file.empty()) {
    std::ifstream f(opt.pos_file);
    if (!f) return fatal("Open position file " + opt.pos_file + " failed!");
    std::string line;
    while (std::getline(f, line)) {
      auto tok = split_ws(line);
      if (tok.size() >= 2) positions.insert(tok[0] + ":" + tok[1]);
    }
  }
  std::map<std::string, int> chrs;
  {
    std::stringstream ss(opt.chrs2process);
    std::string c;
    while (std::getline(ss, c, ',')) if (!c.empty()) chrs[c]++;
  }

  std::vector<std::mutex> ctx_lock((size_t)n_gpu);

  const double t_buf0 = wall();
  const int np = ped.n_person();
  size_t batch = opt.batch_sites > 0 ? (size_t)opt.batch_sites : (size_t)1 << 16;
  // keep a batch of packed input below ~256 MB
  while (batch > 1024 && batch * (size_t)np * sizeof(pm_person_site) > ((size_t)256 << 20)) batch >>= 1;
  // batch buffers: page-locked when the engine offers it, so H2D/D2H overlap the kernels
  struct HostBuf {
    const Engine &e; void *p = nullptr;
    HostBuf(const Engine &eng, size_t bytes) : e(eng) { p = e.host_alloc ? e.host_alloc(bytes) : calloc(bytes ? bytes : 1, 1); }  // (every byte the engine or the writer reads is written first)
    ~HostBuf() { if (e.host_free) e.host_free(p); else free(p); }
    HostBuf(const HostBuf &) = delete;
  };
  struct Slot {
    std::unique_ptr<HostBuf> b_hdr, b_ps, b_status, b_res, b_pres;
    pm_site_hdr *hdr; pm_person_site *ps; uint16_t *status; pm_site_result *res; pm_person_result *pres;
    size_t n = 0, n_res = 0;
    size_t cap = 0;               // rows b_res / b_pres can hold
    std::future<int> fut;
    std::string err;
    std::string text;             // the batch's VCF rows, formatted on the worker thread
    std::vector<size_t> row_end;  // text offset after each row
  };
  const size_t n_slots = (size_t)(n_gpu == 1 ? 2 : 2 * n_gpu);
  std::vector<Slot> slots(n_slots);
  // Row buffers: every site can be a row under --all_sites / --pos; otherwise rows are rare, so start small and let a
  // batch that overflows (PM_EINVAL with the needed count) grow its slot and run again.
  const size_t cap0 = (opt.out_all_sites || opt.force_call) ? batch : std::max<size_t>(64, batch / 16);
  auto size_rows = [&](Slot &sl, size_t cap) -> bool {
    sl.b_res.reset(); sl.b_pres.reset();
    sl.b_res.reset(new HostBuf(engine, cap * sizeof(pm_site_result)));
    sl.b_pres.reset(new HostBuf(engine, cap * (size_t)np * sizeof(pm_person_result)));
    if (!sl.b_res->p || !sl.b_pres->p) return false;
    sl.res = (pm_site_result *)sl.b_res->p; sl.pres = (pm_person_result *)sl.b_pres->p;
    sl.cap = cap;
    return true;
  };
  for (Slot &sl : slots) {
    sl.b_hdr.reset(new HostBuf(engine, batch * sizeof(pm_site_hdr)));
    sl.b_ps.reset(new HostBuf(engine, batch * (size_t)np * sizeof(pm_person_site)));
    sl.b_status.reset(new HostBuf(engine, batch * sizeof(uint16_t)));
    if (!sl.b_hdr->p || !sl.b_ps->p || !sl.b_status->p || !size_rows(sl, cap0)) { destroy_all(); return fatal("out of host memory for the site batches"); }
    sl.hdr = (pm_site_hdr *)sl.b_hdr->p; sl.ps = (pm_person_site *)sl.b_ps->p; sl.status = (uint16_t *)sl.b_status->p;
  }
  // rows of a batch are formatted on the batch's worker plus helpers: --ingest_threads (default: all cores) over the slots
  int fmt_threads = opt.ingest_threads > 0 ? opt.ingest_threads : (int)std::thread::hardware_concurrency();
  fmt_threads = std::max(1, std::min(32, fmt_threads) / (int)n_slots);
  const bool timing = getenv("PM_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_ingest = 0.0, t_wait = 0.0, t_write = 0.0;
  const double t_setup_done = now();

  VcfWriter writer(vcf, opt, ped);
  time_t t0;
  time(&t0);
  printf("Analysis started on %s\n", ctime(&t0));
  size_t out_cnt = 0;
  bool postprob_ran = false;  // has any row's genotype posteriors been computed yet in this run (see consume())
  pm_site_result first_res;
  std::vector<pm_person_result> first_pres((size_t)np);
  int processed_chrs = 0;
  bool stop = false;
  std::string engine_error;
  try {
    while (!stop && glf.next_section()) {
      if (!chrs.empty() && processed_chrs >= (int)chrs.size()) break;
      const std::string label = glf.label();
      if (!chrs.empty() && chrs.count(label) == 0) continue;  // next_section() skips what is left of this one
      uint8_t chr_class = label == opt.chrX ? PM_CHR_X : label == opt.chrY ? PM_CHR_Y : label == opt.chrMT ? PM_CHR_MT : PM_CHR_AUTO;
      Counters cnt;
      processed_chrs++;
      time_t tc;
      time(&tc);
      std::deque<size_t> in_flight;  // slot indices, oldest first
      size_t launched = 0;
      // waits for the oldest batch, updates the counters and prints its rows (in site order)
      auto consume = [&]() -> bool {
        Slot &sl = slots[in_flight.front()];
        in_flight.pop_front();
        const double tw0 = now();
        const int rc = sl.fut.get();
        t_wait += now() - tw0;
        if (rc != PM_OK) { engine_error = sl.err; return false; }
        if (stop) return true;  // --pos already satisfied: drain without printing
        if (!postprob_ran && !sl.row_end.empty()) {
          // The first genotype posteriors of the run.  On chrX / chrY / MT without --denovo the reference's nuclear
          // code sees its initial `sex` member there (PM_HDR_FIRST_POSTPROB in the header): that one site again.
          postprob_ran = true;
          if (!opt.denovo && chr_class != PM_CHR_AUTO) {
            const size_t s0 = sl.res[0].site;
            pm_site_hdr h1 = sl.hdr[s0];
            h1.reserved |= PM_HDR_FIRST_POSTPROB;
            uint16_t st1 = 0;
            size_t n1 = 0;
            int rc1;
            {
              std::lock_guard<std::mutex> guard(ctx_lock[0]);
              rc1 = engine.call_glf(ctxs[0], &h1, &sl.ps[s0 * (size_t)np], 1, &st1, &first_res, first_pres.data(), 1, &n1);
              if (rc1 != PM_OK) engine_error = engine.last_error();
            }
            if (rc1 != PM_OK) return false;
            if (n1 != 1 || (st1 & 0xf) != PM_SITE_EMITTED) throw std::runtime_error("first-row recomputation did not emit the site");
            first_res.site = (uint32_t)s0;
            sl.res[0] = first_res;
            memcpy(sl.pres, first_pres.data(), sizeof(pm_person_result) * (size_t)np);
            std::string row;
            writer.format_site(row, label, sl.hdr[s0], sl.res[0], &sl.ps[s0 * (size_t)np], sl.pres);
            const size_t old = sl.row_end[0];
            sl.text.replace(0, old, row);
            for (size_t &e : sl.row_end) e = e - old + row.size();
          }
        }
        for (size_t s = 0; s < sl.n; s++) count_site(cnt, sl.hdr[s], sl.status[s], opt);
        // rows and dropped de novo candidates in site order: the first of either prints the header
        size_t next_row = 0;
        for (size_t s = 0; s < sl.n && !stop; s++) {
          const int code = sl.status[s] & 0xf;
          if (code == PM_SITE_DENOVO_DROPPED) { writer.ensure_header(); continue; }
          if (code != PM_SITE_EMITTED) continue;
          next_row++;
          out_cnt++;
          if (opt.force_call && out_cnt >= positions.size()) stop = true;  // main.cpp:593
        }
        if (next_row > sl.row_end.size()) throw std::runtime_error("engine returned rows out of site order");
        const double tw1 = now();
        if (next_row > 0) writer.write_rows(sl.text.data(), sl.row_end[next_row - 1], (long)next_row);
        t_write += now() - tw1;
        return true;
      };
      bool more = true, ok = true;
      while (more && !stop && ok) {
        if (in_flight.size() == n_slots) ok = consume();
        if (!ok || stop) break;
        const size_t si = launched % n_slots;
        Slot &sl = slots[si];
        const double ti0 = now();
        size_t n = glf.next_batch(sl.hdr, sl.ps, batch);
        t_ingest += now() - ti0;
        if (n == 0) { more = false; break; }
        if (cnt.totalEntryCnt == 0) cnt.totalEntryCnt = glf.max_position();
        for (size_t s = 0; s < n; s++) sl.hdr[s].chr_class = chr_class;
        if (!positions.empty()) {  // --pos: keep the listed positions only (main.cpp:332-337)
          size_t k = 0;
          for (size_t s = 0; s < n; s++) {
            if (positions.count(label + ":" + std::to_string(sl.hdr[s].pos + 1)) == 0) continue;
            if (k != s) { sl.hdr[k] = sl.hdr[s]; memcpy(&sl.ps[k * (size_t)np], &sl.ps[s * (size_t)np], sizeof(pm_person_site) * (size_t)np); }
            k++;
          }
          n = k;
          if (n == 0) continue;
        }
        sl.n = n; sl.n_res = 0; sl.err.clear();
        const size_t g = launched % (size_t)n_gpu;
        Slot *slp = &sl;
        sl.fut = std::async(std::launch::async, [&, slp, g, label]() -> int {
          std::lock_guard<std::mutex> guard(ctx_lock[g]);
          int rc = engine.call_glf(ctxs[g], slp->hdr, slp->ps, slp->n, slp->status, slp->res, slp->pres, slp->cap, &slp->n_res);
          if (rc == PM_EINVAL && slp->n_res > slp->cap && slp->n_res <= slp->n) {  // more rows than the slot holds: grow, run again
            if (!size_rows(*slp, std::min(batch, slp->n_res + slp->n_res / 4 + 64))) { slp->err = "out of host memory for the result rows"; return PM_EINVAL; }
            rc = engine.call_glf(ctxs[g], slp->hdr, slp->ps, slp->n, slp->status, slp->res, slp->pres, slp->cap, &slp->n_res);
          }
          if (rc != PM_OK) { slp->err = engine.last_error(); return rc; }  // the message is thread-local: keep it
          // the batch's rows as text, here on the worker so that formatting overlaps the next batch's GPU time
          slp->text.clear(); slp->row_end.clear();
          std::vector<size_t> row_site;
          for (size_t s = 0; s < slp->n; s++) {
            if ((slp->status[s] & 0xf) != PM_SITE_EMITTED) continue;
            const size_t r = row_site.size();
            if (r >= slp->n_res || slp->res[r].site != s) { slp->err = "engine returned rows out of site order"; return PM_EINVAL; }
            row_site.push_back(s);
          }
          const size_t nr = row_site.size();
          const size_t parts = nr >= 64 ? std::min<size_t>((size_t)fmt_threads, nr / 32) : 1;
          std::vector<std::string> part_text(parts);
          std::vector<std::vector<size_t>> part_end(parts);
          auto format_part = [&](size_t k) {
            for (size_t r = nr * k / parts; r < nr * (k + 1) / parts; r++) {
              const size_t s = row_site[r];
              writer.format_site(part_text[k], label, slp->hdr[s], slp->res[r], &slp->ps[s * (size_t)np], &slp->pres[r * (size_t)np]);
              part_end[k].push_back(part_text[k].size());
            }
          };
          {
            std::vector<std::thread> pool;
            for (size_t k = 1; k < parts; k++) pool.emplace_back(format_part, k);
            format_part(0);
            for (auto &th : pool) th.join();
          }
          if (parts == 1) slp->text.swap(part_text[0]);
          else {
            size_t total = 0;
            for (auto &t : part_text) total += t.size();
            slp->text.reserve(total);
          }
          for (size_t k = 0, off = 0; k < parts; k++) {
            if (parts > 1) slp->text += part_text[k];
            for (size_t e : part_end[k]) slp->row_end.push_back(off + e);
            off = parts > 1 ? slp->text.size() : 0;
          }
          return rc;
        });
        in_flight.push_back(si);
        launched++;
      }
      while (!in_flight.empty()) ok = consume() && ok;
      if (!ok) {
        destroy_all();
        fclose(vcf);
        return fatal(std::string("engine '") + engine.name + "': " + engine_error);
      }
      if (stop) break;
      // summary block, src/main.cpp:596-621
      int totalBases = 0;
      for (int i = 0; i < 5; i++) totalBases += cnt.refBaseCounts[i];
      int other = cnt.tstvs1 + cnt.tstvs2 + cnt.tvs1tvs2;
      printf("Summary of reference -- %s\n", label.c_str());
      printf("Total Entry Count: %9d\n", cnt.totalEntryCnt);
      printf("Total Base Cout: %9d\n", totalBases);
      printf("Non-Polymorphic Count: %9d\n", cnt.homoRef);
      printf("Transition Count: %9d\n", cnt.transitions);
      printf("Transversion Count: %9d\n", cnt.transversions);
      printf("Other Polymorphism Count: %9d\n", other);
      printf("Filter counts:\n");
      printf("\tminMapQual %u\n", cnt.minMapQualFilter);
      printf("\tminTotalDepth %u\n", cnt.minTotalDepthFilter);
      printf("\tmaxTotalDepth %u\n", cnt.maxTotalDepthFilter);
      printf("Hard to call: %9d\n", cnt.nocall);
      printf("Skipped bases: %u\n", (unsigned)(cnt.totalEntryCnt - cnt.homoRef - cnt.transitions - cnt.transversions - other));
      time_t t1;
      time(&t1);
      printf("Analysis ended on %s\n", ctime(&t1));
      printf("Running time is %u seconds\n\n", (unsigned)(t1 - tc));
      fflush(vcf);
    }
  } catch (const std::exception &e) {
    for (Slot &sl : slots) if (sl.fut.valid()) sl.fut.wait();
    destroy_all();
    fclose(vcf);
    return fatal(e.what());
  }
  if (timing)
    fprintf(stderr, "[pm timing] setup %.3f s (flags + pedigree %.3f, GLF open %.3f alongside %d engine context(s) %.3f, %zu x %zu-site batch buffers %.3f); "
                    "loop %.3f s = ingest %.3f + waiting for the engine/formatting %.3f + writing %.3f\n",
            t_setup_done - t_start, t_ped_done - t_start, t_open_done - t_ped_done, n_gpu, t_ctx, n_slots, batch, t_setup_done - t_buf0,
            now() - t_setup_done, t_ingest, t_wait, t_write);
  destroy_all();
  fclose(vcf);
  return 0;
}

}  // namespace pmh
