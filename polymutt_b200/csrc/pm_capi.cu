// pm_capi.cu — the C ABI (include/polymutt_b200.h) over the sm_100a kernels: context set-up
// (pedigree -> device descriptors, host-computed tables), device scratch management, the host-buffer
// and device-buffer calling entry points, timing and the roofline microbenchmarks.
//
// There is no CPU path here: every entry point that computes needs a CUDA device and fails with
// PM_ECUDA otherwise.
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "host/host_error.h"
#include "pm_device.cuh"
#include "pm_kernels.h"

using pmh::fail;

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) return fail(PM_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct pm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  pm_params par{};
  int n_person = 0, n_fam = 0, n_units = 0, n_es = 0;
  pm::LaunchPlan plan{};
  // device descriptors
  pm::DevRun *d_run = nullptr;
  pm::DevFam *d_fams = nullptr;
  pm::DevUnit *d_units = nullptr;
  uint8_t *d_sex = nullptr;
  // --quick_call pre-pass (main.cpp:354-437): the same kernels on a second description of the run in which everybody
  // is an unrelated founder; its per-site verdicts overrule the real pass (k_quick_merge)
  pm::LaunchPlan plan_q{};
  pm::DevRun *d_run_q = nullptr;
  pm::DevUnit *d_units_q = nullptr;
  pm_site_result *d_res_q = nullptr;
  uint16_t *d_status_q = nullptr;
  // VCF input: a second description in which every family with non-founders is peeled, for chrX / chrY / MT records
  // (FamilyLikelihoodSeq_VCF.cpp:101, 148); each description only does the records of its chromosome class
  bool have_x = false;
  bool batch_has_nonauto = false;
  pm::LaunchPlan plan_x{};
  pm::DevRun *d_run_x = nullptr;
  pm::DevFam *d_fams_x = nullptr;
  pm::DevUnit *d_units_x = nullptr;
  int32_t *d_es_x = nullptr;
  pm::DevStep *d_steps_x = nullptr;
  int32_t *d_es = nullptr;
  pm::DevStep *d_steps = nullptr;
  int *d_err = nullptr;
  double *d_spill = nullptr;  // wide kernel: coefficients of the units beyond threads * units_per_thread (shared by the passes of a batch: stream-ordered)
  size_t cap_spill = 0;
  unsigned long long *d_counters = nullptr;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;
  // scratch (grown on demand)
  size_t cap_sites = 0;
  pm_site_result *d_res_all = nullptr;
  uint32_t *d_emit_sites = nullptr;
  uint32_t *d_tile_scratch = nullptr;   // k_compact_*: emitted sites per tile of 1,024
  uint32_t *d_n_emit = nullptr;
  // staging for the host-buffer entry point
  size_t cap_in_sites = 0, cap_out_rows = 0;
  double *d_mono[2] = {nullptr, nullptr}; size_t cap_mono = 0;
  pm_site_hdr *d_hdr[2] = {nullptr, nullptr};   // two input slots: H2D of chunk k+1 overlaps compute of chunk k
  uint4 *d_recs[2] = {nullptr, nullptr};
  // VCF input (vcf_records_chunked): 3-byte PL triplets as they arrive, two output slots, a stream for the results
  uint8_t *d_pl3[2] = {nullptr, nullptr}; size_t cap_pl3 = 0;
  pm_site_result *d_vres[2] = {nullptr, nullptr}; size_t cap_vres = 0;
  pm_person_result *d_vperson[2] = {nullptr, nullptr}; size_t cap_vperson = 0;
  uint16_t *d_vcalls[2] = {nullptr, nullptr}; size_t cap_vcalls = 0;
  cudaStream_t stream_d2h = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  unsigned char *d_wire[2] = {nullptr, nullptr}; size_t cap_wire_sites = 0;   // 14-byte records as they arrive (pm_call_glf_sites_wire)
  cudaStream_t stream_h2d = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
  uint32_t *h_rows = nullptr;                   // pinned
  uint16_t *d_status = nullptr;                 // VCF input: the per-record status words (not copied back)
  // GLF host path: two output slots (status, compacted rows, row count)
  uint16_t *d_gstatus[2] = {nullptr, nullptr};
  pm_site_result *d_gres[2] = {nullptr, nullptr};
  pm_person_result *d_gperson[2] = {nullptr, nullptr};
  uint32_t *d_gemit[2] = {nullptr, nullptr};
  struct RowFix { size_t first, rows; uint32_t base; };
  std::vector<RowFix> row_fix;
  // timing of the last call
  float ms_main = 0.f, ms_total = 0.f;
  int launches = 0;
  bool timing_cached = false;  // set by the host-buffer entry point (sums over its chunks)
};

namespace {

template <typename T>
int dev_alloc(T **p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e != cudaSuccess) { *p = nullptr; return fail(e == cudaErrorMemoryAllocation ? PM_ENOMEM : PM_ECUDA, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); }
  return PM_OK;
}

int ensure_spill(pm_ctx *c) {
  size_t need = 0;
  for (const pm::LaunchPlan *p : {&c->plan, &c->plan_q, &c->plan_x})
    if (p->kind == pm::LaunchPlan::WIDE) need = std::max(need, (size_t)p->grid * (size_t)p->n_spill * 5);
  if (need <= c->cap_spill) return PM_OK;
  c->cap_spill = 0;
  int rc = dev_alloc(&c->d_spill, need);
  if (rc) return rc;
  c->cap_spill = need;
  return PM_OK;
}

int ensure_scratch(pm_ctx *c, size_t n_sites) {
  if (n_sites <= c->cap_sites) return PM_OK;
  int rc;
  c->cap_sites = 0;
  if ((rc = dev_alloc(&c->d_res_all, n_sites))) return rc;
  if ((rc = dev_alloc(&c->d_emit_sites, n_sites))) return rc;
  if ((rc = dev_alloc(&c->d_tile_scratch, (n_sites + 1023) / 1024))) return rc;
  if (c->par.quick_call) {
    if ((rc = dev_alloc(&c->d_res_q, n_sites))) return rc;
    if ((rc = dev_alloc(&c->d_status_q, n_sites))) return rc;
  }
  c->cap_sites = n_sites;
  return PM_OK;
}

}  // namespace

extern "C" pm_ctx *pm_create(const pm_pedigree *ped, const pm_params *par, const double *lut256, int device) {
  pmh::clear_error();
  if (!ped || !par || ped->n_fam <= 0 || ped->n_person <= 0) { fail(PM_EINVAL, "pm_create: empty pedigree or missing parameters"); return nullptr; }
  if (par->quick_call && par->vcf_input) { fail(PM_EINVAL, "pm_create: --quick_call does not exist for VCF input"); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { fail(PM_ECUDA, "pm_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e)); return nullptr; }
  if (device < 0 || device >= ndev) { fail(PM_EINVAL, "pm_create: device %d out of range (0..%d)", device, ndev - 1); return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { fail(PM_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); return nullptr; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { fail(PM_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { fail(PM_ECUDA, "pm_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return nullptr; }

  pm_ctx *c = new pm_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->par = *par;
  c->n_person = ped->n_person;
  c->n_fam = ped->n_fam;

  // ---- host-side descriptors ----
  // all_peeled: every family with non-founders is an extended family (VCF input on chrX / chrY / MT, FLSeq_VCF.cpp:101)
  struct Desc {
    std::vector<pm::DevFam> fams;
    std::vector<pm::DevUnit> units;
    std::vector<int32_t> es;
    std::vector<pm::DevStep> steps;
    int founders_total = 0, kids_total = 0;
  };
  auto build_desc = [&](Desc &D, bool all_peeled) -> bool {
  std::vector<pm::DevFam> &fams = D.fams;
  std::vector<pm::DevUnit> &units = D.units;
  std::vector<int32_t> &es = D.es;
  std::vector<pm::DevStep> &steps = D.steps;
  int &founders_total = D.founders_total, &kids_total = D.kids_total;
  fams.assign((size_t)ped->n_fam, pm::DevFam());
  int first = 0;
  for (int f = 0; f < ped->n_fam; f++) {
    pm::DevFam &d = fams[(size_t)f];
    memset(&d, 0, sizeof d);
    const int size = ped->fam_size[f], nf = ped->fam_founders[f];
    d.first = first; d.size = (int16_t)size; d.founders = (int16_t)nf;
    founders_total += nf;
    // VCF mode uses the nuclear-family formula only when there are several families (FamilyLikelihoodSeq_VCF.cpp:98-103);
    // a lone nuclear family goes through the bi-allelic Elston-Stewart peel there.
    const bool nuclear = ped->fam_generations[f] == 2 && nf == 2 && !(par->vcf_input && ped->n_fam == 1) && !all_peeled;
    if (size == nf) {
      d.kind = 0;
      for (int j = 0; j < size; j++) units.push_back({first + j, -1, kids_total, (int32_t)ped->sex[first + j]});
    } else if (nuclear) {
      d.kind = 1;
      units.push_back({first, size - 2, kids_total, 0});
      kids_total += size - 2;
    } else {
      d.kind = 2;
      if (size > pm::kMaxEsPersons) { fail(PM_EUNSUPPORTED, "extended family %d has %d members; the device peel workspace holds %d", f, size, pm::kMaxEsPersons); return false; }
      if (!ped->peel || !ped->peel_first) { fail(PM_EINVAL, "pm_create: extended family %d but no peeling order was supplied", f); return false; }
      const int p0 = ped->peel_first[f], p1 = ped->peel_first[f + 1];
      if (p1 <= p0) { fail(PM_EINVAL, "pm_create: extended family %d has an empty peeling order", f); return false; }
      d.step_first = (int32_t)steps.size(); d.n_steps = (int16_t)(p1 - p0);
      // resolve the std::map<pair,...> marriage_partials lookups of the reference (ES:1084-1087,
      // 1147, 1236) once: exact (first, second) key match.
      std::map<std::pair<int, int>, int> slots;
      for (int s = p0; s < p1; s++) {
        const pm_peel_step &ps = ped->peel[s];
        pm::DevStep ds;
        memset(&ds, 0, sizeof ds);
        ds.type = (int8_t)ps.type; ds.from0 = (int8_t)ps.from0; ds.from1 = (int8_t)ps.from1;
        ds.to0 = (int8_t)ps.to0; ds.to1 = (int8_t)ps.to1; ds.mp = -1;
        if (ps.type == PM_PEEL_CHILD_TO_PARENTS) {
          auto key = std::make_pair(ps.to0, ps.to1);
          auto it = slots.find(key);
          if (it == slots.end()) { int id = (int)slots.size(); slots[key] = id; ds.mp = (int8_t)id; ds.flag = 1; }
          else ds.mp = (int8_t)it->second;
        } else if (ps.type == PM_PEEL_SPOUSE_TO_SPOUSE) {
          const bool from_is_mother = ped->sex[first + ps.from0] == 2;  // ES:1137-1148
          auto key = from_is_mother ? std::make_pair(ps.to0, ps.from0) : std::make_pair(ps.from0, ps.to0);
          ds.flag = from_is_mother ? 0 : 1;
          auto it = slots.find(key);
          if (it != slots.end()) ds.mp = (int8_t)it->second;
        } else if (ps.type == PM_PEEL_PARENTS_TO_CHILD) {
          auto it = slots.find(std::make_pair(ps.from0, ps.from1));
          if (it != slots.end()) ds.mp = (int8_t)it->second;
        } else {
          fail(PM_EINVAL, "pm_create: bad peeling step type %d", ps.type); return false;
        }
        steps.push_back(ds);
      }
      if ((int)slots.size() > pm::kMaxMp) { fail(PM_EUNSUPPORTED, "extended family %d needs %zu marriage partials; the device workspace holds %d", f, slots.size(), pm::kMaxMp); return false; }
      d.n_mp = (int8_t)slots.size();
      es.push_back(f);
    }
    first += size;
  }
  if (first != ped->n_person) { fail(PM_EINVAL, "pm_create: n_person != sum(fam_size)"); return false; }
  if (founders_total == 0) { fail(PM_EINVAL, "Family size is zero"); return false; }
  return true;
  };
  Desc D0;
  if (!build_desc(D0, false)) { delete c; return nullptr; }
  std::vector<pm::DevFam> &fams = D0.fams;
  std::vector<pm::DevUnit> &units = D0.units;
  std::vector<int32_t> &es = D0.es;
  std::vector<pm::DevStep> &steps = D0.steps;
  const int founders_total = D0.founders_total, kids_total = D0.kids_total;
  c->n_units = (int)units.size();
  c->n_es = (int)es.size();

  pm::DevRun run;
  memset(&run, 0, sizeof run);
  if (lut256) memcpy(run.lut, lut256, sizeof run.lut); else pm_fill_lut(run.lut);
  pm_genotype_mutation_matrix(par->denovo_mut_rate, par->denovo_tstv, run.mut);
  // SetPolyPrior (NucFam:231-242) and the per-hypothesis prior terms of main:447-533
  double prior = 0;
  for (int i = 1; i <= 2 * founders_total; i++) prior += 1.0 / i;
  prior *= par->theta;
  const double prior_ts = par->poly_tstv / (par->poly_tstv + 1), prior_tv = (1 - prior_ts) / 2;
  run.log_1m_prior = log10(1 - prior);
  run.log_prior_ts = log10(prior * prior_ts);
  run.log_prior_tv = log10(prior * prior_tv);
  run.log_prior_other = log10(prior * 0.001);
  run.log_prior_23 = log10(prior * 2. / 3.);
  run.log_prior_16 = log10(prior * 1. / 6.);
  {  // SetPolyPrior_chrX / _chrY / _MT (NucFam:256-293): founder chromosomes by sex
    int male_founders = 0, female_founders = 0;
    for (int i = 0; i < ped->n_person; i++)
      if (ped->father[i] < 0 && ped->mother[i] < 0) { male_founders += ped->sex[i] == 1; female_founders += ped->sex[i] == 2; }
    const int n_chr[4] = {2 * founders_total, 2 * female_founders + male_founders, male_founders, founders_total};
    for (int cl = 0; cl < 4; cl++) {
      double pr = 0;
      for (int i = 1; i <= n_chr[cl]; i++) pr += 1.0 / i;
      pr *= par->theta;
      run.cls_log[cl][0] = log10(1 - pr);
      run.cls_log[cl][1] = log10(pr * prior_ts);
      run.cls_log[cl][2] = log10(pr * prior_tv);
      run.cls_log[cl][3] = log10(pr * 0.001);
      run.cls_log[cl][4] = log10(pr * 2. / 3.);
      run.cls_log[cl][5] = log10(pr * 1. / 6.);
    }
  }
  run.log_min_llr = log10(par->denovo_min_llr);
  run.theta = par->theta; run.posterior_cutoff = par->posterior_cutoff; run.precision = par->precision;
  run.denovo_min_llr = par->denovo_min_llr; run.min_ps = par->min_ps;
  run.min_map_quality = par->min_map_quality; run.min_total_depth = par->min_total_depth; run.max_total_depth = par->max_total_depth;
  run.denovo = par->denovo; run.force_call = par->force_call; run.out_all_sites = par->out_all_sites;
  run.n_person = ped->n_person; run.n_fam = ped->n_fam; run.n_units = c->n_units; run.n_es = c->n_es; run.n_kids = kids_total;
  for (int i = 0; i < 128; i++) {
    run.log_inv[i] = 1.0 / (1.0 + (i + 0.5) / 128.0);
    run.log_tab[i] = (double)(-log10l((long double)run.log_inv[i]));
  }
  run.use_brent = (ped->n_fam > 1 || fams[0].kind != 1 || par->vcf_input) ? 1 : 0;  // FLSeq:94; always in VCF mode
  run.vcf_mode = par->vcf_input ? 1 : 0;
  run.site_filter = par->vcf_input ? 1 : 0;
  // PedVCF::tstv_ratio is hard-wired to 2.0 (PedVCF.cpp:7); GetPolyPrior_indel returns the SNP prior (NucFam:313)
  run.vcf_log_ts = log10(2.0 / (2.0 + 1));
  run.vcf_log_tv = log10(0.5 / (2.0 + 1));
  run.vcf_log_indel = log10(prior);

  e = pm::plan_launch(&c->plan, c->n_person, c->n_units, c->n_es, c->sm_count, nullptr);
  c->plan.ten_state = par->denovo != 0;  // (DevRun::denovo, which the kernels read)
  if (e == cudaErrorNotSupported) {
    fail(PM_EUNSUPPORTED, "pedigree shape not supported by the device kernels yet (%d quartic units, %d extended families, %d persons)", c->n_units, c->n_es, c->n_person);
    delete c; return nullptr;
  }
  if (e != cudaSuccess) { fail(PM_ECUDA, "kernel set-up: %s", cudaGetErrorString(e)); delete c; return nullptr; }

  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_h2d[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_h2d[1], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess && cudaEventCreate(&c->ev2) == cudaSuccess;
  ok = ok && dev_alloc(&c->d_run, 1) == PM_OK && dev_alloc(&c->d_fams, fams.size()) == PM_OK &&
       dev_alloc(&c->d_units, units.size()) == PM_OK && dev_alloc(&c->d_es, es.size()) == PM_OK &&
       dev_alloc(&c->d_steps, steps.size()) == PM_OK && dev_alloc(&c->d_sex, (size_t)ped->n_person) == PM_OK && dev_alloc(&c->d_err, 2) == PM_OK && dev_alloc(&c->d_n_emit, 1) == PM_OK &&
       dev_alloc(&c->d_counters, 16) == PM_OK && cudaEventCreate(&c->tm0) == cudaSuccess && cudaEventCreate(&c->tm1) == cudaSuccess;
  if (ok) {
    run.fams = c->d_fams; run.units = c->d_units; run.es_fams = c->d_es; run.steps = c->d_steps;
    run.counters = c->d_counters;
    run.sex = c->d_sex;
    ok = cudaMemcpy(c->d_sex, ped->sex, (size_t)ped->n_person, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c->d_fams, fams.data(), fams.size() * sizeof(pm::DevFam), cudaMemcpyHostToDevice) == cudaSuccess &&
         (units.empty() || cudaMemcpy(c->d_units, units.data(), units.size() * sizeof(pm::DevUnit), cudaMemcpyHostToDevice) == cudaSuccess) &&
         (es.empty() || cudaMemcpy(c->d_es, es.data(), es.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess) &&
         (steps.empty() || cudaMemcpy(c->d_steps, steps.data(), steps.size() * sizeof(pm::DevStep), cudaMemcpyHostToDevice) == cudaSuccess) &&
         cudaMemcpy(c->d_run, &run, sizeof run, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemset(c->d_err, 0, 2 * sizeof(int)) == cudaSuccess && cudaMemset(c->d_counters, 0, 16 * sizeof(unsigned long long)) == cudaSuccess;
  }
  if (ok && par->vcf_input) {
    Desc D1;
    if (!build_desc(D1, true)) { pm_destroy(c); return nullptr; }
    pm::DevRun run_x = run;
    run_x.n_units = (int)D1.units.size(); run_x.n_es = (int)D1.es.size(); run_x.n_kids = 0; run_x.site_filter = 2;
    cudaError_t ex = pm::plan_launch(&c->plan_x, c->n_person, run_x.n_units, run_x.n_es, c->sm_count, nullptr);
    c->plan_x.ten_state = run_x.denovo != 0;
    if (ex == cudaSuccess) {
      ok = dev_alloc(&c->d_run_x, 1) == PM_OK && dev_alloc(&c->d_fams_x, D1.fams.size()) == PM_OK && dev_alloc(&c->d_units_x, D1.units.size()) == PM_OK &&
           dev_alloc(&c->d_es_x, D1.es.size()) == PM_OK && dev_alloc(&c->d_steps_x, D1.steps.size()) == PM_OK;
      if (ok) {
        run_x.fams = c->d_fams_x; run_x.units = c->d_units_x; run_x.es_fams = c->d_es_x; run_x.steps = c->d_steps_x;
        ok = cudaMemcpy(c->d_fams_x, D1.fams.data(), D1.fams.size() * sizeof(pm::DevFam), cudaMemcpyHostToDevice) == cudaSuccess &&
             (D1.units.empty() || cudaMemcpy(c->d_units_x, D1.units.data(), D1.units.size() * sizeof(pm::DevUnit), cudaMemcpyHostToDevice) == cudaSuccess) &&
             (D1.es.empty() || cudaMemcpy(c->d_es_x, D1.es.data(), D1.es.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess) &&
             (D1.steps.empty() || cudaMemcpy(c->d_steps_x, D1.steps.data(), D1.steps.size() * sizeof(pm::DevStep), cudaMemcpyHostToDevice) == cudaSuccess) &&
             cudaMemcpy(c->d_run_x, &run_x, sizeof run_x, cudaMemcpyHostToDevice) == cudaSuccess;
        c->have_x = ok;
      }
    }  // else: chrX / chrY / MT records are refused when they come (pm_call_vcf_records)
  }
  if (ok && par->quick_call) {
    // MakeUnrelated() (FLSeq.cpp:55-60): founders = count in every family, so every person is a single founder, no
    // family isNuclear() and PolymorphismLogLikelihood always runs Brent (FLSeq.cpp:94).  The pre-pass never uses the
    // --denovo model for H0 (main.cpp:370) and stops at the `continue`s of main.cpp:432-433.
    std::vector<pm::DevUnit> units_q;
    for (int i = 0; i < ped->n_person; i++) units_q.push_back({i, -1, 0, (int32_t)ped->sex[i]});
    pm::DevRun run_q = run;
    run_q.n_units = ped->n_person; run_q.n_es = 0; run_q.n_kids = 0; run_q.use_brent = 1;
    run_q.denovo = 0; run_q.force_call = 0; run_q.out_all_sites = 0;
    run_q.counters = c->d_counters + 8;  // its work is counted apart
    e = pm::plan_launch(&c->plan_q, c->n_person, ped->n_person, 0, c->sm_count, nullptr);
    if (e != cudaSuccess) {
      fail(e == cudaErrorNotSupported ? PM_EUNSUPPORTED : PM_ECUDA, "--quick_call: the unrelated pre-pass needs %d single-founder units (%s)", ped->n_person, cudaGetErrorString(e));
      pm_destroy(c);
      return nullptr;
    }
    ok = dev_alloc(&c->d_run_q, 1) == PM_OK && dev_alloc(&c->d_units_q, units_q.size()) == PM_OK;
    if (ok) {
      run_q.units = c->d_units_q;
      ok = cudaMemcpy(c->d_units_q, units_q.data(), units_q.size() * sizeof(pm::DevUnit), cudaMemcpyHostToDevice) == cudaSuccess &&
           cudaMemcpy(c->d_run_q, &run_q, sizeof run_q, cudaMemcpyHostToDevice) == cudaSuccess;
    }
  }
  if (ok && ensure_spill(c) != PM_OK) ok = false;
  if (!ok) {
    if (!*pmh::last_error()) fail(PM_ECUDA, "pm_create: device set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    pm_destroy(c);
    return nullptr;
  }
  return c;
}

extern "C" void pm_destroy(pm_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->stream_h2d) cudaStreamSynchronize(c->stream_h2d);  // a copy issued before an error return may still be reading the caller's buffer
  if (c->stream_d2h) cudaStreamSynchronize(c->stream_d2h);
  cudaFree(c->d_run_x); cudaFree(c->d_fams_x); cudaFree(c->d_units_x); cudaFree(c->d_es_x); cudaFree(c->d_steps_x);
  cudaFree(c->d_sex); cudaFree(c->d_run_q); cudaFree(c->d_units_q); cudaFree(c->d_res_q); cudaFree(c->d_status_q);
  cudaFree(c->d_run); cudaFree(c->d_fams); cudaFree(c->d_units); cudaFree(c->d_es); cudaFree(c->d_steps);
  cudaFree(c->d_err); cudaFree(c->d_spill); cudaFree(c->d_counters); cudaFree(c->d_res_all); cudaFree(c->d_emit_sites); cudaFree(c->d_tile_scratch); cudaFree(c->d_n_emit);
  for (int k = 0; k < 2; k++) { cudaFree(c->d_hdr[k]); cudaFree(c->d_recs[k]); cudaFree(c->d_wire[k]); cudaFree(c->d_pl3[k]); cudaFree(c->d_vres[k]); cudaFree(c->d_vperson[k]); cudaFree(c->d_vcalls[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
    if (c->ev_d2h[k]) cudaEventDestroy(c->ev_d2h[k]);
    if (c->ev_h2d[k]) cudaEventDestroy(c->ev_h2d[k]); }
  if (c->stream_h2d) cudaStreamDestroy(c->stream_h2d);
  if (c->stream_d2h) cudaStreamDestroy(c->stream_d2h);
  if (c->h_rows) cudaFreeHost(c->h_rows);
  cudaFree(c->d_status);
  for (int k = 0; k < 2; k++) { cudaFree(c->d_gstatus[k]); cudaFree(c->d_gres[k]); cudaFree(c->d_gperson[k]); cudaFree(c->d_gemit[k]); }
  cudaFree(c->d_mono[0]); cudaFree(c->d_mono[1]);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->tm0) cudaEventDestroy(c->tm0);
  if (c->tm1) cudaEventDestroy(c->tm1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static int run_device(pm_ctx *c, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                      size_t n_sites, int out_mode, uint16_t *d_status_out, pm_site_result *d_res_out,
                      pm_person_result *d_person_out, size_t res_cap, uint32_t *d_n_res, uint16_t *d_calls_out = nullptr);
static int check_device_error(pm_ctx *c);

// the result stream and the per-slot events of the host-buffer entry points, made on first use
static int ensure_copy_streams(pm_ctx *c) {
  if (c->stream_d2h) return PM_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_d2h[k], cudaEventDisableTiming));
  }
  return PM_OK;
}

extern "C" int pm_call_glf_sites_device(pm_ctx *c, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site,
                                        size_t n_sites, int out_mode, uint16_t *d_status_out, pm_site_result *d_res_out,
                                        pm_person_result *d_person_out, size_t res_cap, uint32_t *d_n_res) {
  if (!c) return fail(PM_EINVAL, "null context");
  if (c->par.vcf_input) return fail(PM_EINVAL, "this ctx was created for VCF input; use pm_call_vcf_records");
  return run_device(c, d_hdr, d_person_site, nullptr, n_sites, out_mode, d_status_out, d_res_out, d_person_out, res_cap, d_n_res);
}

// d_calls_out != nullptr: the per-person results leave as 2 bytes (best | gq << 8) there and d_person_out is not used
static int run_device(pm_ctx *c, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                      size_t n_sites, int out_mode, uint16_t *d_status_out, pm_site_result *d_res_out,
                      pm_person_result *d_person_out, size_t res_cap, uint32_t *d_n_res, uint16_t *d_calls_out) {
  if (n_sites == 0) { if (d_n_res) CUDA_TRY(cudaMemsetAsync(d_n_res, 0, sizeof(uint32_t), c->stream)); return PM_OK; }
  if (!d_hdr || !d_person_site || !d_status_out || !d_res_out || (!d_person_out && !d_calls_out))
    return fail(PM_EINVAL, "pm_call_glf_sites_device: null buffer");
  if (n_sites > 0xffffffffull) return fail(PM_EINVAL, "at most 2^32-1 sites per call");
  if (out_mode == PM_OUT_ALL && res_cap < n_sites) return fail(PM_EINVAL, "PM_OUT_ALL needs res_cap >= n_sites");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_scratch(c, n_sites);
  if (rc) return rc;
  uint32_t *d_cnt = d_n_res ? d_n_res : c->d_n_emit;
  CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  if (c->par.quick_call)
    CUDA_TRY(pm::launch_sites(c->plan_q, c->d_run_q, d_hdr, (const uint4 *)d_person_site, nullptr, n_sites, c->d_spill, c->d_res_q, c->d_status_q, c->d_err, c->stream));
  CUDA_TRY(pm::launch_sites(c->plan, c->d_run, d_hdr, (const uint4 *)d_person_site, d_mono, n_sites, c->d_spill, c->d_res_all, d_status_out, c->d_err, c->stream));
  if (c->par.quick_call) CUDA_TRY(pm::launch_quick_merge(c->d_status_q, n_sites, c->d_res_all, d_status_out, c->stream));
  const bool second = c->par.vcf_input && c->have_x && c->batch_has_nonauto;
  if (second)
    CUDA_TRY(pm::launch_sites(c->plan_x, c->d_run_x, d_hdr, (const uint4 *)d_person_site, d_mono, n_sites, c->d_spill, c->d_res_all, d_status_out, c->d_err, c->stream));
  CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  const bool ten_state = c->par.denovo && !c->par.vcf_input;
  const bool with_ab = !c->par.denovo && !c->par.vcf_input;  // the allele balance is printed by the non-de-novo GLF writer only
  CUDA_TRY(pm::launch_compact(d_status_out, n_sites, c->d_emit_sites, d_cnt, out_mode == PM_OUT_ALL, c->d_tile_scratch, c->stream));
  CUDA_TRY(pm::launch_post(c->d_run, c->n_fam, d_hdr, (const uint4 *)d_person_site, c->d_res_all, c->d_emit_sites, d_cnt,
                           out_mode == PM_OUT_ALL ? n_sites : (res_cap < n_sites ? res_cap : n_sites), res_cap, d_res_out,
                           d_person_out, d_calls_out, c->n_es > 0, ten_state, c->sm_count, with_ab, c->stream));
  if (second)
    CUDA_TRY(pm::launch_post(c->d_run_x, c->n_fam, d_hdr, (const uint4 *)d_person_site, c->d_res_all, c->d_emit_sites, d_cnt,
                             out_mode == PM_OUT_ALL ? n_sites : (res_cap < n_sites ? res_cap : n_sites), res_cap, d_res_out,
                             d_person_out, d_calls_out, true, ten_state, c->sm_count, false, c->stream));
  CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
  // the site kernel's autosomal instance + its (normally empty) chrX/Y/MT one, k_compact (three launches for long batches) or
  // k_all_rows, k_post [, k_post_es] [, k_post_ab]
  c->launches = 2 + ((out_mode != PM_OUT_ALL && n_sites > 8 * 1024) ? 3 : 1) + 1 + (c->n_es > 0 ? 1 : 0) + (with_ab ? 1 : 0);
  if (c->par.quick_call) c->launches += 3;
  if (second) c->launches += 4;  // both instances of the site kernel, k_post, k_post_es on the all-families-peeled description
  c->timing_cached = false;
  return PM_OK;
}

extern "C" int pm_sync(pm_ctx *c) {
  if (!c) return fail(PM_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return check_device_error(c);
}

extern "C" int pm_last_timing(pm_ctx *c, float *ms_main_kernel, float *ms_total, int *n_launches) {
  if (!c) return fail(PM_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  float a = 0.f, b = 0.f;
  if (c->timing_cached) {
    a = c->ms_main; b = c->ms_total;
  } else {
    CUDA_TRY(cudaEventSynchronize(c->ev2));
    CUDA_TRY(cudaEventElapsedTime(&a, c->ev0, c->ev1));
    CUDA_TRY(cudaEventElapsedTime(&b, c->ev0, c->ev2));
  }
  if (ms_main_kernel) *ms_main_kernel = a;
  if (ms_total) *ms_total = b;
  if (n_launches) *n_launches = c->launches;
  return PM_OK;
}

// Host-buffer entry point.  The batch is cut into chunks of ~48 MB of packed input that go through two input slots and
// two output slots on three streams: while chunk k is computed, chunk k+1 is copied in and the rows of chunk k-1 are
// copied out -- the host only waits for a chunk's row count (it sizes the copy of its rows) after the next chunk's
// kernels are already queued, so the call runs at the slower of PCIe and the kernels.  Pinned host buffers
// (pm_host_alloc) make the copies truly asynchronous; pageable buffers work too, just without the overlap.
// rec_bytes: 16 (pm_person_site) or 14 (pm_person_site_wire: copied into a staging slot and widened on the device by
// k_unpack_wire, on the compute stream right in front of the chunk's kernels).
static int glf_sites_host(pm_ctx *c, const pm_site_hdr *hdr, const void *person_site, size_t rec_bytes, size_t n_sites,
                          int out_mode, uint16_t *status_out, pm_site_result *res_out, pm_person_result *person_out,
                          size_t res_cap, size_t *n_res) {
  if (!c) return fail(PM_EINVAL, "null context");
  if (n_res) *n_res = 0;
  if (n_sites == 0) return PM_OK;
  if (c->par.vcf_input) return fail(PM_EINVAL, "this ctx was created for VCF input; use pm_call_vcf_records");
  if (!hdr || !person_site || !res_out) return fail(PM_EINVAL, "pm_call_glf_sites: null buffer");
  for (size_t s = 0; s < n_sites; s++)
    if (hdr[s].chr_class > PM_CHR_MT) return fail(PM_EINVAL, "site %zu: chr_class %d is not one of PM_CHR_*", s, (int)hdr[s].chr_class);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t np = (size_t)c->n_person;
  size_t chunk = ((size_t)48 << 20) / (np * sizeof(pm_person_site));
  if (c->plan.kind == pm::LaunchPlan::NARROW) {
    // One thread per site: a launch lasts a whole number of "waves" of resident threads (2 to 5 blocks of 128 per SM by
    // instance), each as long as one site's serial evaluation (~1 ms for a 20-member peel).  A 48 MB chunk of a 20-person
    // pedigree is 157 k sites = 2.07 waves of 75.8 k: it paid for 3.  Four times the largest resident set is a whole
    // number of waves for every instance (up to 256 MB of records).
    const size_t waves4 = (size_t)4 * (size_t)c->sm_count * 5 * 128;
    const size_t cap = ((size_t)256 << 20) / (np * sizeof(pm_person_site));
    chunk = std::max(chunk, std::min(waves4, cap));
  }
  if (chunk < 256) chunk = 256;
  if (chunk > ((size_t)1 << 20)) chunk = (size_t)1 << 20;
  if (chunk > n_sites) chunk = n_sites;
  int rc;
  if (chunk > c->cap_in_sites) {
    c->cap_in_sites = 0;  // a failed allocation below must not leave a stale capacity behind
    for (int k = 0; k < 2; k++) {
      if ((rc = dev_alloc(&c->d_hdr[k], chunk))) return rc;
      if ((rc = dev_alloc(&c->d_recs[k], chunk * np))) return rc;
    }
    c->cap_in_sites = chunk;
  }
  if (chunk > c->cap_out_rows) {  // a chunk can emit at most `chunk` rows
    c->cap_out_rows = 0;
    for (int k = 0; k < 2; k++) {
      if ((rc = dev_alloc(&c->d_gstatus[k], chunk))) return rc;
      if ((rc = dev_alloc(&c->d_gres[k], chunk))) return rc;
      if ((rc = dev_alloc(&c->d_gperson[k], chunk * np))) return rc;
      if ((rc = dev_alloc(&c->d_gemit[k], 1))) return rc;
    }
    c->cap_out_rows = chunk;
  }
  const bool wire = rec_bytes != sizeof(pm_person_site);
  if (wire && chunk > c->cap_wire_sites) {
    c->cap_wire_sites = 0;
    for (int k = 0; k < 2; k++)
      if ((rc = dev_alloc(&c->d_wire[k], chunk * np * rec_bytes + 16))) return rc;  // + 16: the unpack kernel reads whole uint4s
    c->cap_wire_sites = chunk;
  }
  if (!c->h_rows) CUDA_TRY(cudaHostAlloc((void **)&c->h_rows, 2 * sizeof(uint32_t), cudaHostAllocDefault));
  if ((rc = ensure_copy_streams(c))) return rc;
  const size_t n_chunks = (n_sites + chunk - 1) / chunk;
  auto chunk_len = [&](size_t k) { return k + 1 < n_chunks ? chunk : n_sites - k * chunk; };
  auto issue_h2d = [&](size_t k) -> cudaError_t {
    const int slot = (int)(k & 1);
    const size_t base = k * chunk, n = chunk_len(k);
    cudaError_t e = cudaSuccess;
    if (k >= 2) e = cudaStreamWaitEvent(c->stream_h2d, c->ev_done[slot], 0);  // the slot's last reader: chunk k-2's kernels
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_hdr[slot], hdr + base, n * sizeof(pm_site_hdr), cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(wire ? (void *)c->d_wire[slot] : (void *)c->d_recs[slot], (const unsigned char *)person_site + base * np * rec_bytes,
                          n * np * rec_bytes, cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_h2d[slot], c->stream_h2d);
    return e;
  };
  // an error return must not leave a copy in flight that still reads (or writes) the caller's buffers
  auto bail = [&](int code) { cudaStreamSynchronize(c->stream_h2d); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream_d2h); return code; };
  size_t total_rows = 0;
  bool overflow = false;
  int launches = 0;
  cudaError_t e = cudaSuccess;
  // chunk k's kernels have been queued: wait for its row count and send its rows home (on the result stream)
  auto finish = [&](size_t k) -> int {
    const int slot = (int)(k & 1);
    const size_t base = k * chunk;
    if ((e = cudaEventSynchronize(c->ev_done[slot])) != cudaSuccess) return fail(PM_ECUDA, "GLF sites chunk %zu: %s", k, cudaGetErrorString(e));
    const uint32_t rows = c->h_rows[slot];
    if (total_rows + rows > res_cap) { overflow = true; total_rows += rows; return PM_OK; }
    if (rows) {
      e = cudaMemcpyAsync(res_out + total_rows, c->d_gres[slot], rows * sizeof(pm_site_result), cudaMemcpyDeviceToHost, c->stream_d2h);
      if (e == cudaSuccess && person_out)
        e = cudaMemcpyAsync(person_out + total_rows * np, c->d_gperson[slot], rows * np * sizeof(pm_person_result), cudaMemcpyDeviceToHost, c->stream_d2h);
      if (e != cudaSuccess) return fail(PM_ECUDA, "GLF sites chunk %zu results: %s", k, cudaGetErrorString(e));
      c->row_fix.push_back({total_rows, (size_t)rows, (uint32_t)base});
    }
    if ((e = cudaEventRecord(c->ev_d2h[slot], c->stream_d2h)) != cudaSuccess) return fail(PM_ECUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
    total_rows += rows;
    return PM_OK;
  };
  c->row_fix.clear();
  if ((e = issue_h2d(0)) != cudaSuccess) return bail(fail(PM_ECUDA, "H2D copy: %s", cudaGetErrorString(e)));
  for (size_t k = 0; k < n_chunks; k++) {
    const int slot = (int)(k & 1);
    const size_t base = k * chunk, n = chunk_len(k);
    if (k + 1 < n_chunks && (e = issue_h2d(k + 1)) != cudaSuccess) return bail(fail(PM_ECUDA, "H2D copy: %s", cudaGetErrorString(e)));
    if ((e = cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0)) != cudaSuccess) return bail(fail(PM_ECUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e)));
    // the output slot was last copied out for chunk k-2 (finish(k-2) recorded the event, rows or not)
    if (k >= 2 && (e = cudaStreamWaitEvent(c->stream, c->ev_d2h[slot], 0)) != cudaSuccess) return bail(fail(PM_ECUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e)));
    if (wire && (e = pm::launch_unpack_wire(c->d_wire[slot], c->d_recs[slot], n * np, c->sm_count, c->stream)) != cudaSuccess)
      return bail(fail(PM_ECUDA, "k_unpack_wire: %s", cudaGetErrorString(e)));
    rc = pm_call_glf_sites_device(c, c->d_hdr[slot], (const pm_person_site *)c->d_recs[slot], n, out_mode, c->d_gstatus[slot], c->d_gres[slot],
                                  c->d_gperson[slot], c->cap_out_rows, c->d_gemit[slot]);
    if (rc) return bail(rc);
    launches += c->launches + (wire ? 1 : 0);
    e = cudaMemcpyAsync(c->h_rows + slot, c->d_gemit[slot], sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && status_out) e = cudaMemcpyAsync(status_out + base, c->d_gstatus[slot], n * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_done[slot], c->stream);
    if (e != cudaSuccess) return bail(fail(PM_ECUDA, "GLF sites chunk %zu: %s", k, cudaGetErrorString(e)));
    if (k >= 1 && (rc = finish(k - 1))) return bail(rc);
  }
  if ((rc = finish(n_chunks - 1))) return bail(rc);
  e = cudaStreamSynchronize(c->stream_d2h);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return bail(fail(PM_ECUDA, "GLF sites: %s", cudaGetErrorString(e)));
  for (const pm_ctx::RowFix &f : c->row_fix)   // site indices of a chunk's rows are chunk-relative on the device
    for (size_t r = 0; r < f.rows; r++) res_out[f.first + r].site += f.base;
  c->launches = launches;
  c->timing_cached = false;  // (the events hold the last chunk's times)
  if (n_res) *n_res = total_rows;
  if ((rc = check_device_error(c))) return rc;  // the error word a kernel sets: polled once per call
  if (overflow) return fail(PM_EINVAL, "pm_call_glf_sites: res_cap %zu too small, %zu rows needed", res_cap, total_rows);
  return PM_OK;
}

extern "C" int pm_call_glf_sites(pm_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, size_t n_sites,
                                 int out_mode, uint16_t *status_out, pm_site_result *res_out, pm_person_result *person_out,
                                 size_t res_cap, size_t *n_res) {
  return glf_sites_host(c, hdr, person_site, sizeof(pm_person_site), n_sites, out_mode, status_out, res_out, person_out, res_cap, n_res);
}

extern "C" int pm_call_glf_sites_wire(pm_ctx *c, const pm_site_hdr *hdr, const pm_person_site_wire *person_site_wire, size_t n_sites,
                                      int out_mode, uint16_t *status_out, pm_site_result *res_out, pm_person_result *person_out,
                                      size_t res_cap, size_t *n_res) {
  static_assert(sizeof(pm_person_site_wire) == 14, "pm_person_site_wire must be 14 bytes");
  return glf_sites_host(c, hdr, person_site_wire, sizeof(pm_person_site_wire), n_sites, out_mode, status_out, res_out, person_out, res_cap, n_res);
}

// Checks the device-side error word (set by a kernel that met a site it cannot handle); one small blocking copy, so once per call.
static int check_device_error(pm_ctx *c) {
  int err = 0;
  CUDA_TRY(cudaMemcpy(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    cudaMemset(c->d_err, 0, sizeof(int));
    if (err == PM_EUNSUPPORTED) return fail(PM_EUNSUPPORTED, "a site's chr_class is not one of PM_CHR_*");
    return fail(err, "device-side error %d", err);
  }
  return PM_OK;
}

// VCF-input records from host buffers, in chunks of ~128 MB of 16-byte records, through two input slots and two output
// slots on three streams: while chunk k is computed, chunk k+1 comes in and chunk k-1's results go out; nothing waits for
// the host inside the loop.  Input: 16-byte records, or (pl3) three PL bytes per sample widened on the device.  Output per
// sample: person_out (96 bytes) or calls_out (2 bytes: best | gq << 8, all the --in_vcf writer prints from, written by
// k_post directly) -- exactly one of the two.
static int vcf_records_chunked(pm_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, const uint8_t *pl3, const double *mono,
                               size_t n, pm_site_result *res_out, pm_person_result *person_out, uint16_t *calls_out) {
  if (!c) return fail(PM_EINVAL, "null context");
  if (!c->par.vcf_input) return fail(PM_EINVAL, "pm_call_vcf_records: the ctx was not created with pm_params.vcf_input = 1");
  if (n == 0) return PM_OK;
  if (!hdr || (!person_site && !pl3) || !mono || !res_out) return fail(PM_EINVAL, "pm_call_vcf_records: null buffer");
  for (size_t s = 0; s < n; s++) {
    const int a2 = hdr[s].reserved & 0xff;
    if (hdr[s].ref_base < 1 || hdr[s].ref_base > 4 || a2 < 1 || a2 > 4 || a2 == hdr[s].ref_base)
      return fail(PM_EINVAL, "record %zu: alleles must be two different bases in 1..4 (got %d, %d)", s, hdr[s].ref_base, a2);
    if (hdr[s].chr_class > PM_CHR_MT) return fail(PM_EINVAL, "record %zu: chr_class %d is not one of PM_CHR_*", s, (int)hdr[s].chr_class);
    if (hdr[s].chr_class != PM_CHR_AUTO && !c->have_x)
      return fail(PM_EUNSUPPORTED, "record %zu: chrX/chrY/MT records need every family peeled, which the device kernels cannot do for this pedigree", s);
  }
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t np = (size_t)c->n_person;
  size_t chunk = ((size_t)128 << 20) / (np * sizeof(pm_person_site));
  if (chunk < 256) chunk = 256;
  if (chunk > ((size_t)1 << 20)) chunk = (size_t)1 << 20;
  if (chunk > n) chunk = n;
  int rc;
  if (chunk > c->cap_in_sites) {
    c->cap_in_sites = 0;
    for (int k = 0; k < 2; k++) {
      if ((rc = dev_alloc(&c->d_hdr[k], chunk))) return rc;
      if ((rc = dev_alloc(&c->d_recs[k], chunk * np))) return rc;
    }
    if ((rc = dev_alloc(&c->d_status, chunk))) return rc;
    c->cap_in_sites = chunk;
  }
  if (pl3 && chunk > c->cap_pl3) {
    c->cap_pl3 = 0;
    for (int k = 0; k < 2; k++) if ((rc = dev_alloc(&c->d_pl3[k], chunk * np * 3))) return rc;
    c->cap_pl3 = chunk;
  }
  if (chunk > c->cap_vres) {
    c->cap_vres = 0;
    for (int k = 0; k < 2; k++) if ((rc = dev_alloc(&c->d_vres[k], chunk))) return rc;
    c->cap_vres = chunk;
  }
  if (person_out && chunk > c->cap_vperson) {
    c->cap_vperson = 0;
    for (int k = 0; k < 2; k++) if ((rc = dev_alloc(&c->d_vperson[k], chunk * np))) return rc;
    c->cap_vperson = chunk;
  }
  if (calls_out && chunk > c->cap_vcalls) {
    c->cap_vcalls = 0;
    for (int k = 0; k < 2; k++) if ((rc = dev_alloc(&c->d_vcalls[k], chunk * np))) return rc;
    c->cap_vcalls = chunk;
  }
  if (chunk > c->cap_mono) {
    c->cap_mono = 0;
    for (int k = 0; k < 2; k++) if ((rc = dev_alloc(&c->d_mono[k], chunk))) return rc;
    c->cap_mono = chunk;
  }
  if ((rc = ensure_copy_streams(c))) return rc;
  const size_t n_chunks = (n + chunk - 1) / chunk;
  auto chunk_len = [&](size_t k) { return k + 1 < n_chunks ? chunk : n - k * chunk; };
  auto issue_h2d = [&](size_t k) -> cudaError_t {
    const int slot = (int)(k & 1);
    const size_t base = k * chunk, m = chunk_len(k);
    cudaError_t e = cudaSuccess;
    if (k >= 2) e = cudaStreamWaitEvent(c->stream_h2d, c->ev_done[slot], 0);  // the slot's last reader: chunk k-2's kernels
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_hdr[slot], hdr + base, m * sizeof(pm_site_hdr), cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess && pl3) e = cudaMemcpyAsync(c->d_pl3[slot], pl3 + base * np * 3, m * np * 3, cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess && !pl3) e = cudaMemcpyAsync(c->d_recs[slot], person_site + base * np, m * np * sizeof(pm_person_site), cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_mono[slot], mono + base, m * sizeof(double), cudaMemcpyHostToDevice, c->stream_h2d);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_h2d[slot], c->stream_h2d);
    return e;
  };
  // nothing may still read or write the caller's buffers when the call returns, error or not
  auto drain = [&]() -> cudaError_t {
    cudaError_t e0 = cudaStreamSynchronize(c->stream_h2d), e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->stream_d2h);
    return e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2);
  };
  auto bail = [&](int code) { drain(); return code; };
  cudaError_t e = issue_h2d(0);
  if (e != cudaSuccess) return bail(fail(PM_ECUDA, "H2D copy: %s", cudaGetErrorString(e)));
  int launches = 0;
  for (size_t k = 0; k < n_chunks; k++) {
    const int slot = (int)(k & 1);
    const size_t base = k * chunk, m = chunk_len(k);
    if (k + 1 < n_chunks && (e = issue_h2d(k + 1)) != cudaSuccess) return bail(fail(PM_ECUDA, "H2D copy: %s", cudaGetErrorString(e)));
    if ((e = cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0)) != cudaSuccess) return bail(fail(PM_ECUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e)));
    // the output slot was last copied out for chunk k-2
    if (k >= 2 && (e = cudaStreamWaitEvent(c->stream, c->ev_d2h[slot], 0)) != cudaSuccess) return bail(fail(PM_ECUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e)));
    c->batch_has_nonauto = false;
    for (size_t r = 0; r < m; r++) c->batch_has_nonauto |= hdr[base + r].chr_class != PM_CHR_AUTO;
    if (pl3 && (e = pm::launch_unpack_pl3(c->d_hdr[slot], c->d_pl3[slot], c->d_recs[slot], m, (int)np, c->sm_count, c->stream)) != cudaSuccess)
      return bail(fail(PM_ECUDA, "k_unpack_pl3: %s", cudaGetErrorString(e)));
    rc = run_device(c, c->d_hdr[slot], (const pm_person_site *)c->d_recs[slot], c->d_mono[slot], m, PM_OUT_ALL, c->d_status, c->d_vres[slot],
                    person_out ? c->d_vperson[slot] : nullptr, m, c->d_n_emit, calls_out ? c->d_vcalls[slot] : nullptr);
    if (rc) return bail(rc);
    launches += c->launches + (pl3 ? 1 : 0);
    e = cudaEventRecord(c->ev_done[slot], c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream_d2h, c->ev_done[slot], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(res_out + base, c->d_vres[slot], m * sizeof(pm_site_result), cudaMemcpyDeviceToHost, c->stream_d2h);
    if (e == cudaSuccess && person_out)
      e = cudaMemcpyAsync(person_out + base * np, c->d_vperson[slot], m * np * sizeof(pm_person_result), cudaMemcpyDeviceToHost, c->stream_d2h);
    if (e == cudaSuccess && calls_out)
      e = cudaMemcpyAsync(calls_out + base * np, c->d_vcalls[slot], m * np * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream_d2h);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_d2h[slot], c->stream_d2h);
    if (e != cudaSuccess) return bail(fail(PM_ECUDA, "VCF records chunk %zu: %s", k, cudaGetErrorString(e)));
  }
  if ((e = drain()) != cudaSuccess) return fail(PM_ECUDA, "VCF records: %s", cudaGetErrorString(e));
  for (size_t k = 1; k < n_chunks; k++)
    for (size_t r = 0, base = k * chunk, m = chunk_len(k); r < m; r++) res_out[base + r].site += (uint32_t)base;
  c->launches = launches;
  c->timing_cached = false;  // (the events hold the last chunk's times)
  return check_device_error(c);
}

extern "C" int pm_call_vcf_records(pm_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono,
                                   size_t n, pm_site_result *res_out, pm_person_result *person_out) {
  if (n && (!person_out || !person_site)) return fail(PM_EINVAL, "pm_call_vcf_records: null buffer");
  return vcf_records_chunked(c, hdr, person_site, nullptr, mono, n, res_out, person_out, nullptr);
}

extern "C" int pm_call_vcf_records_calls(pm_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono,
                                         size_t n, pm_site_result *res_out, uint16_t *calls_out) {
  if (n && (!calls_out || !person_site)) return fail(PM_EINVAL, "pm_call_vcf_records_calls: null buffer");
  return vcf_records_chunked(c, hdr, person_site, nullptr, mono, n, res_out, nullptr, calls_out);
}

extern "C" int pm_call_vcf_records_pl(pm_ctx *c, const pm_site_hdr *hdr, const uint8_t *pl3, const double *mono, size_t n,
                                      pm_site_result *res_out, uint16_t *calls_out) {
  if (n && (!calls_out || !pl3)) return fail(PM_EINVAL, "pm_call_vcf_records_pl: null buffer");
  return vcf_records_chunked(c, hdr, nullptr, pl3, mono, n, res_out, nullptr, calls_out);
}

// Device-buffer variant of pm_call_vcf_records (bench.py's device-resident leg): every record gets a row.
extern "C" int pm_call_vcf_records_device(pm_ctx *c, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                                          size_t n, int has_nonauto, uint16_t *d_status_out, pm_site_result *d_res_out,
                                          pm_person_result *d_person_out) {
  if (!c) return fail(PM_EINVAL, "null context");
  if (!c->par.vcf_input) return fail(PM_EINVAL, "pm_call_vcf_records_device: the ctx was not created with pm_params.vcf_input = 1");
  if (has_nonauto && !c->have_x) return fail(PM_EUNSUPPORTED, "chrX/chrY/MT records need every family peeled, which the device kernels cannot do for this pedigree");
  if (n && !d_mono) return fail(PM_EINVAL, "pm_call_vcf_records_device: null buffer");
  c->batch_has_nonauto = has_nonauto != 0;
  return run_device(c, d_hdr, d_person_site, d_mono, n, PM_OUT_ALL, d_status_out, d_res_out, d_person_out, n, c->d_n_emit);
}

// The same with 2 bytes per sample out (d_calls_out[n * n_person] = best | gq << 8) instead of the 96-byte rows.
extern "C" int pm_call_vcf_records_calls_device(pm_ctx *c, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                                                size_t n, int has_nonauto, uint16_t *d_status_out, pm_site_result *d_res_out,
                                                uint16_t *d_calls_out) {
  if (!c) return fail(PM_EINVAL, "null context");
  if (!c->par.vcf_input) return fail(PM_EINVAL, "pm_call_vcf_records_calls_device: the ctx was not created with pm_params.vcf_input = 1");
  if (has_nonauto && !c->have_x) return fail(PM_EUNSUPPORTED, "chrX/chrY/MT records need every family peeled, which the device kernels cannot do for this pedigree");
  if (n && (!d_mono || !d_calls_out)) return fail(PM_EINVAL, "pm_call_vcf_records_calls_device: null buffer");
  c->batch_has_nonauto = has_nonauto != 0;
  return run_device(c, d_hdr, d_person_site, d_mono, n, PM_OUT_ALL, d_status_out, d_res_out, nullptr, n, c->d_n_emit, d_calls_out);
}

extern "C" void *pm_host_alloc(size_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { fail(PM_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
extern "C" void pm_host_free(void *p) { if (p) cudaFreeHost(p); }

// Human-readable description of the kernel plan
extern "C" int pm_describe_plan(pm_ctx *c, char *buf, size_t len) {
  if (!c || !buf || !len) return fail(PM_EINVAL, "null argument");
  if (c->plan.kind == pm::LaunchPlan::NARROW) snprintf(buf, len, "k_sites_narrow<%d>: one thread per site, %d threads/block", c->plan.units_per_thread, c->plan.threads);
  else snprintf(buf, len, "k_sites_wide<U=%d>: one block of %d threads per site, %d units/thread in registers (%d units, %d in the L2 scratch), persistent grid %d (%d blocks/SM), "
                "%s, %s",
                c->plan.units_per_thread, c->plan.threads, c->plan.units_per_thread, c->n_units, c->plan.n_spill, c->plan.grid, c->plan.blocks_per_sm,
                c->plan.variant == 6 ? "the site's records read from global memory (too large for shared memory)" : "one TMA bulk copy per site",
                c->plan.f3_offset && !c->plan.n_spill && !c->plan.es ? "H1-H3 of an autosomal site in one pass over Brent's monotone path (3 block barriers per monomorphic site)"
                                                                      : "one hypothesis at a time over Brent's monotone path");
  return PM_OK;
}

// Test / tuning hook: re-plans the main pass on a given instantiation of the wide kernel (pm_wide.cu: PM_WIDE_VARIANTS) with
// `threads` threads per block, also for pedigrees the narrow kernel would normally take.  Extended families need the ES instances,
// which every variant has.
extern "C" int pm_force_wide_plan(pm_ctx *c, int variant, int threads) {
  if (!c) return fail(PM_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int force[2] = {variant, threads};
  pm::LaunchPlan p{};
  cudaError_t e = pm::plan_launch(&p, c->n_person, c->n_units, c->n_es, c->sm_count, force);
  if (e == cudaErrorInvalidValue) return fail(PM_EINVAL, "pm_force_wide_plan: no variant %d with %d threads", variant, threads);
  if (e == cudaErrorNotSupported) return fail(PM_EUNSUPPORTED, "pm_force_wide_plan: variant %d with %d threads cannot hold this pedigree", variant, threads);
  if (e != cudaSuccess) return fail(PM_ECUDA, "pm_force_wide_plan: %s", cudaGetErrorString(e));
  p.ten_state = c->par.denovo != 0;
  c->plan = p;
  return ensure_spill(c);
}

extern "C" int pm_timer_start(pm_ctx *c) {
  if (!c) return fail(PM_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventRecord(c->tm0, c->stream));
  return PM_OK;
}
extern "C" int pm_timer_stop(pm_ctx *c, float *ms) {
  if (!c || !ms) return fail(PM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventRecord(c->tm1, c->stream));
  CUDA_TRY(cudaEventSynchronize(c->tm1));
  CUDA_TRY(cudaEventElapsedTime(ms, c->tm0, c->tm1));
  return PM_OK;
}
extern "C" int pm_get_counters(pm_ctx *c, pm_counters *out) {
  if (!c || !out) return fail(PM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(out, c->d_counters, sizeof *out, cudaMemcpyDeviceToHost));
  return PM_OK;
}
// Debug hook (PM_PHASE_TIMING builds only): cycles thread 0 of the wide kernel spent per phase, summed over sites:
// [0] TMA wait, [1] stats, [2] set-up, [3] evaluation (to the first barrier), [4] serial tail, [5] decisions/refit/write.
extern "C" int pm_debug_phase_cycles(pm_ctx *c, unsigned long long *out8) {
  if (!c || !out8) return fail(PM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(out8, c->d_counters + 8, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return PM_OK;
}

extern "C" int pm_reset_counters(pm_ctx *c) {
  if (!c) return fail(PM_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemsetAsync(c->d_counters, 0, 16 * sizeof(unsigned long long), c->stream));
  return PM_OK;
}

extern "C" int pm_measure_fp64_peak(pm_ctx *c, double *flops) {
  if (!c || !flops) return fail(PM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  const int blocks = c->sm_count * 8, threads = 256, iters = 20000;
  double *d_out = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d_out, (size_t)blocks * threads * sizeof(double)));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(pm::launch_dfma_peak(d_out, blocks, threads, iters, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaFree(d_out);
  *flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads / (best * 1e-3);
  return PM_OK;
}

extern "C" int pm_measure_copy_bw(pm_ctx *c, double *bytes_per_s) {
  if (!c || !bytes_per_s) return fail(PM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t bytes = (size_t)1 << 30;
  void *a = nullptr, *b = nullptr;
  CUDA_TRY(cudaMalloc(&a, bytes));
  cudaError_t e = cudaMalloc(&b, bytes);
  if (e != cudaSuccess) { cudaFree(a); return fail(PM_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  cudaMemsetAsync(a, 1, bytes, c->stream);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(pm::launch_copy(a, b, bytes, c->sm_count, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaFree(a); cudaFree(b);
  *bytes_per_s = 2.0 * (double)bytes / (best * 1e-3);
  return PM_OK;
}
