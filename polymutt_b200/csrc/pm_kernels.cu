// pm_kernels.cu — the site kernels (sm_100a).
//
//  k_sites_narrow : one THREAD per site.  Small pedigrees (<= kNarrowMaxUnits quartic units plus any
//                   number of extended families peeled by Elston–Stewart).  Sites of adjacent lanes are
//                   adjacent in HBM, so a warp streams 32 * n_person * 16 contiguous bytes.
//  k_sites_wide   : one BLOCK per site, pedigrees with many units: pm_wide.cu.
//  k_compact      : ordered compaction of the emitted sites (single block scan, deterministic).
//  k_post         : genotype posteriors / GQ / dosage / AB for emitted sites, one thread per (site, family): pm_post.cu.
//
// All arithmetic is FP64; inputs are uint8 phred likelihoods.  See DESIGN.md for the data layout and
// the roofline of each kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>

#include "pm_device.cuh"
#include "pm_es.cuh"
#include "pm_kernels.h"
#include "pm_site_logic.cuh"

namespace pm {

// ================================================================================================
// narrow kernel: one thread per site
// ================================================================================================
constexpr int kNarrowThreads = 128;

struct NarrowSmem {
  SmemTables t;
};

// NA = the instance for chrX / chrY / MT sites (see k_sites_wide): the autosomal one has none of those rules compiled in.
// ES = the pedigree has extended families (the peel's registers are only paid for where there is something to peel)
// DN = false: the instance for runs without --denovo of pedigrees with extended families -- the ten-state peel and its
// workspace (the bulk of the stack frame) are not compiled in
template <int UMAX, bool NA, bool ES, bool DN>
struct NarrowEval {
  const DevRun *run;
  const uint4 *recs;  // this site's records
  const NarrowSmem *sm;
  double B[UMAX > 0 ? UMAX : 1][5];  // UMAX = 0: the instance for pedigrees made of extended families only (no quartic unit)
  double C0[9];  // single-nuclear-family mode keeps the nine conditionals
  int g11, g12, g22;
  bool denovo;
  int cls = PM_CHR_AUTO;  // chromosome class of the site; the hypothesis objects' stale `sex` member is 0 (see pm_device.cuh)
  int n_hyp = 0, n_eval = 0;

  __device__ void setup(int a1, int a2, bool dn) {
    g11 = geno_index(a1, a1); g12 = geno_index(a1, a2); g22 = geno_index(a2, a2);
    denovo = dn;
    if (UMAX > 0 && !run->use_brent) {
      const DevUnit u = run->units[0];
      if (!NA || denovo) unit_conditionals(recs, u.first, u.nkids, g11, g12, g22, denovo, sm->t.lut, sm->t.mut, C0);
      else unit_conditionals_nonauto(recs, u.first, u.nkids, g11, g12, g22, cls, 0, sm->t.lut, C0);
      return;
    }
    if constexpr (!NA) {
#pragma unroll
      for (int u = 0; u < UMAX; u++)
        if (u < run->n_units) unit_quartic(recs, run->units[u], g11, g12, g22, denovo, sm->t.lut, sm->t.mut, B[u]);
    } else {
      for (int u = 0; u < UMAX; u++)
        if (u < run->n_units) unit_quartic_nonauto(recs, run->units[u], g11, g12, g22, denovo, cls, 0, sm->t.lut, sm->t.mut, B[u]);
    }
  }
  // sum_f log10 L_f(p), FLSeq:222-240
  __device__ double loglik(double p) const {
    double sum = 0.0;
    const Monomials m = monomials(p);
#pragma unroll
    for (int u = 0; u < UMAX; u++)
      if (u < run->n_units) sum += log10(quartic_eval(B[u], m));
    if constexpr (ES)
    for (int e = 0; e < run->n_es; e++) {
      const DevFam f = run->fams[run->es_fams[e]];
      double lk;
      if constexpr (DN) {
        // under --denovo the three-state peel only serves the rare bi-allelic refit of a called site: the small eager form,
        // which takes no registers from the ten-state peel around it
        lk = denovo ? es_likelihood_impl<10, NA>(run, f, recs, g11, g12, g22, true, p, sm->t.lut, sm->t.mut, -1, -1, cls)
                    : es_likelihood3_eager<NA>(run, f, recs, g11, g12, g22, false, p, sm->t.lut, sm->t.mut, -1, -1, cls);
      } else {
        lk = es_likelihood_impl<3, NA>(run, f, recs, g11, g12, g22, false, p, sm->t.lut, sm->t.mut, -1, -1, cls);
      }
      sum += log10(lk);
    }
    return sum;
  }
  // single nuclear family: fixed parent-pair table (NucFam:383-420), or HW at freq == 1 under --denovo
  __device__ double loglik_fixed(bool hw_at_one) const {
    double pp[9];
    if (hw_at_one) parent_priors(1.0, pp); else single_trio_priors(pp);
    double sum = 0.0;
    for (int j = 0; j < 9; j++) sum += C0[j] * pp[j];
    return log10(sum);
  }
  // PolymorphismLogLikelihood, FLSeq:91-104
  __device__ double optimize(int a1, int a2, bool dn, double *freq) {
    setup(a1, a2, dn);
    n_hyp++;
    if (UMAX > 0 && !run->use_brent) { n_eval++; return loglik_fixed(false); }
    BrentState st;
    brent_begin(st);
    double f;
    do { f = -loglik(st.u); n_eval++; } while (brent_feed(st, f, run->precision));
    *freq = st.min;
    return -st.fmin;
  }
};

#ifndef PM_NARROW_ES3_MINB
#define PM_NARROW_ES3_MINB 4   // CEPH bi-allelic: 1 (158 registers) 114.7, 4 (128) 118.0, 5 (96, spills) 117.9 M sites/s
#endif
template <int UMAX, bool NA, bool ES, bool DN>
__global__ void __launch_bounds__(kNarrowThreads, (ES && !DN && UMAX == 0) ? PM_NARROW_ES3_MINB : 1) k_sites_narrow(const DevRun *__restrict__ run,
                                                                  const pm_site_hdr *__restrict__ hdr,
                                                                  const uint4 *__restrict__ recs_all,
                                                                  const double *__restrict__ mono_all, size_t n_sites,
                                                                  pm_site_result *__restrict__ res,
                                                                  uint16_t *__restrict__ status, int *__restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NarrowSmem *sm = reinterpret_cast<NarrowSmem *>(smem_raw);
  // which instance does what: see k_sites_wide (err[1] = the autosomal instance met a chrX / chrY / MT site)
  if (NA ? ((err[1] == 0 && run->site_filter != 2) || run->site_filter == 1) : run->site_filter == 2) return;
  load_tables(run, &sm->t);
  __syncthreads();
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sites) return;
  const int np = run->n_person;
  const uint4 *recs = recs_all + s * (size_t)np;
  const pm_site_hdr h = hdr[s];

  pm_site_result r;
  memset(&r, 0, sizeof r);
  r.site = (uint32_t)s;
  r.maxidx = -1;
  const int ref = h.ref_base;
  if (ref < 1 || ref > 4) { r.status = PM_SITE_BAD_REF; res[s] = r; status[s] = status_word(r); return; }
  if (h.chr_class > PM_CHR_MT) { atomicExch(err, PM_EUNSUPPORTED); r.status = PM_SITE_BAD_REF; res[s] = r; status[s] = status_word(r); return; }
  if (NA ? h.chr_class == PM_CHR_AUTO : h.chr_class != PM_CHR_AUTO) {  // the other instance's site
    if (!NA) atomicExch(err + 1, 1);
    return;
  }
  const int cls = NA ? h.chr_class : PM_CHR_AUTO;
  const double log_1m_prior = run->cls_log[cls][0];
  if (run->vcf_mode) {  // one record of a VCF: mono is given, one Brent run for (REF, ALT)
    const int a2 = h.reserved & 0xff;
    NarrowEval<UMAX, NA, ES, DN> ev;
    ev.run = run; ev.recs = recs; ev.sm = sm; ev.cls = cls;
    double freq = 0.0;
    const double poly = ev.optimize(ref, a2, false, &freq);
    vcf_record_result(run, r, ref, a2, (h.reserved & 0x100) != 0, mono_all[s], poly, freq);
    res[s] = r;
    status[s] = status_word(r);
    return;
  }

  // CalcReadStats + filters (NucFam:520-546, main:343-348) and MonomorphismLogLikelihood (NucFam:502-517)
  int total_depth = 0, ns = 0, mapq_sum = 0;
  double lk_mono = 0.0;
  const int grr = geno_index(ref, ref);
  for (int i = 0; i < np; i++) {
    uint4 rec = recs[i];
    int d = rec_depth(rec);
    total_depth += d; mapq_sum += rec_mapq(rec); ns += d > 0;
    lk_mono += -(double)rec_lk(rec, grr) / 10;
  }
  r.total_depth = total_depth; r.num_samp = ns;
  if (ns > 0) { r.avg_map_qual = (double)mapq_sum / (double)ns; r.perc_samp = (double)ns / (double)np; }
  if (total_depth < run->min_total_depth) r.status = PM_SITE_MIN_DEPTH;
  else if (run->max_total_depth > 0 && total_depth > run->max_total_depth) r.status = PM_SITE_MAX_DEPTH;
  else if (r.perc_samp * 100 < run->min_ps) r.status = PM_SITE_MIN_PS;
  else if (r.avg_map_qual < run->min_map_quality) r.status = PM_SITE_MIN_MAPQ;
  if (r.status != 0) { res[s] = r; status[s] = status_word(r); return; }

  NarrowEval<UMAX, NA, ES, DN> ev;
  ev.run = run; ev.recs = recs; ev.sm = sm; ev.cls = cls;
  r.reserved = (uint16_t)ref;
  // H0 (main:447-462)
  if (!run->denovo) {
    r.varllk[0] = log_1m_prior + lk_mono;
  } else {
    int a1, a2;
    hyp_alleles(0, ref, a1, a2);
    ev.setup(a1, a2, true);
    double l0 = (UMAX == 0 || run->use_brent) ? ev.loglik(1.0) : ev.loglik_fixed(true);
    r.varllk[0] = log_1m_prior + l0;
  }
  r.varllk_noprior[0] = r.varllk[0] - log_1m_prior;
  r.varfreq[0] = 1.0;
  for (int hix = 1; hix <= 3; hix++) {
    int a1, a2;
    hyp_alleles(hix, ref, a1, a2);
    double freq = 0.0;
    double ml = ev.optimize(a1, a2, run->denovo != 0, &freq);
    site_store_hyp(run, r, hix, ml, freq, cls);
  }
  var_posterior(r, ref, 4);
  if (r.var_post_prob < 0.99) {  // main:499-537
    for (int hix = 4; hix <= 6; hix++) {
      int a1, a2;
      hyp_alleles(hix, ref, a1, a2);
      double freq = 0.0;
      double ml = ev.optimize(a1, a2, run->denovo != 0, &freq);
      site_store_hyp(run, r, hix, ml, freq, cls);
    }
    var_posterior(r, ref, 7);
  }
  if (site_decide(run, r, lk_mono)) {
    double freq = 0.0;
    double lk_poly = ev.optimize(r.allele1, r.allele2, false, &freq);
    site_finish_refit(run, r, lk_poly, freq);
  }
  if (r.status == PM_SITE_EMITTED && run->denovo && r.denovo_lr < run->denovo_min_llr) { r.flags |= PM_FLAG_ROW_DROPPED; r.status = PM_SITE_DENOVO_DROPPED; }
  r.reserved = 0;
  res[s] = r;
  status[s] = status_word(r);
  // work counters: one atomic per warp (redux over the lanes that got here)
  {
    unsigned h = ev.n_hyp + (run->denovo ? 1 : 0), e = ev.n_eval + (run->denovo ? 1 : 0), em = r.status == PM_SITE_EMITTED;
    const unsigned mask = __activemask();
    h = __reduce_add_sync(mask, h); e = __reduce_add_sync(mask, e); em = __reduce_add_sync(mask, em);
    if ((int)(threadIdx.x & 31) == __ffs(mask) - 1) {
      atomicAdd(&run->counters[0], (unsigned long long)h); atomicAdd(&run->counters[1], (unsigned long long)e);
      atomicAdd(&run->counters[2], (unsigned long long)__popc(mask)); atomicAdd(&run->counters[3], (unsigned long long)em);
    }
  }
}

// ================================================================================================
// --quick_call (main.cpp:354-437): the verdict of the everybody-unrelated pre-pass overrules the real pass
// ================================================================================================
__global__ void k_quick_merge(const uint16_t *__restrict__ status_q, size_t n_sites, pm_site_result *__restrict__ res,
                              uint16_t *__restrict__ status) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sites) return;
  const int code_q = status_q[s] & 0xf;
  if (code_q != PM_SITE_NOCALL && code_q != PM_SITE_MONO) return;  // called a variant, or filtered out in both passes alike
  const pm_site_result old = res[s];
  pm_site_result r;
  memset(&r, 0, sizeof r);
  r.site = old.site; r.maxidx = -1; r.status = PM_SITE_QUICK_SKIP;
  r.total_depth = old.total_depth; r.num_samp = old.num_samp; r.perc_samp = old.perc_samp; r.avg_map_qual = old.avg_map_qual;
  res[s] = r;
  status[s] = PM_SITE_QUICK_SKIP;
}

// ================================================================================================
// ordered compaction of emitted sites (deterministic, single block)
// ================================================================================================
__global__ void __launch_bounds__(1024) k_compact(const uint16_t *__restrict__ status, size_t n_sites,
                                                  uint32_t *__restrict__ emit_sites, uint32_t *__restrict__ n_emit,
                                                  int all) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (size_t start = 0; start < n_sites; start += 1024) {
    size_t s = start + tid;
    uint32_t flag = (s < n_sites) && (all || (status[s] & 0xf) == PM_SITE_EMITTED);
    uint32_t x = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    uint32_t off = base + (warp ? warp_tot[warp - 1] : 0) + x - flag;
    if (flag) emit_sites[off] = (uint32_t)s;
    __syncthreads();
    if (tid == 0) base += warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) *n_emit = base;
}

// The same in three small launches for long batches (the single block takes ~1.3 us per 1,024 sites: 2.6 ms of a 23 ms
// step at 2 Mi sites): emitted sites per tile of 1,024 -> exclusive scan of the tile counts (one block) -> scatter.
__global__ void __launch_bounds__(1024) k_compact_count(const uint16_t *__restrict__ status, size_t n_sites, uint32_t *__restrict__ tile_count) {
  const size_t s = (size_t)blockIdx.x * 1024 + threadIdx.x;
  const int flag = s < n_sites && (status[s] & 0xf) == PM_SITE_EMITTED;
  const int c = __syncthreads_count(flag);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(1024) k_compact_scan(uint32_t *__restrict__ tile_count, unsigned n_tiles, uint32_t *__restrict__ n_emit) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (unsigned start = 0; start < n_tiles; start += 1024) {
    const unsigned i = start + tid;
    const uint32_t v = i < n_tiles ? tile_count[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    if (i < n_tiles) tile_count[i] = base + (warp ? warp_tot[warp - 1] : 0) + x - v;  // exclusive prefix: the tile's first row
    __syncthreads();
    if (tid == 0) base += warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) *n_emit = base;
}
__global__ void __launch_bounds__(1024) k_compact_scatter(const uint16_t *__restrict__ status, size_t n_sites, const uint32_t *__restrict__ tile_first,
                                                          uint32_t *__restrict__ emit_sites) {
  __shared__ uint32_t warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t s = (size_t)blockIdx.x * 1024 + tid;
  const uint32_t flag = s < n_sites && (status[s] & 0xf) == PM_SITE_EMITTED;
  const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) warp_tot[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
    warp_tot[lane] = w;  // inclusive
  }
  __syncthreads();
  if (flag) emit_sites[tile_first[blockIdx.x] + (warp ? warp_tot[warp - 1] : 0) + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)s;
}

// PM_OUT_ALL (parity tests, VCF input: every record gets a row): row i is site i, no scan needed
__global__ void k_all_rows(size_t n_sites, uint32_t *__restrict__ emit_sites, uint32_t *__restrict__ n_emit) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sites; i += (size_t)gridDim.x * blockDim.x) emit_sites[i] = (uint32_t)i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_emit = (uint32_t)n_sites;
}

// ================================================================================================
// 14-byte wire records (pm_person_site_wire) -> 16-byte pm_person_site records the site kernels load as one uint4.
// HBM-bound byte shuffling: 14 bytes read + 16 written per record.  A block moves tiles of 1,024 records: 896 coalesced
// 16-byte loads into shared memory, then thread t builds records t, t+256, ... from 32-bit words (record r starts at
// byte 14 r: word 7 (r/2), plus half a word for odd r) and stores them as coalesced uint4s.
// ================================================================================================
constexpr int kUnpackThreads = 256;
constexpr int kUnpackTile = 1024;                       // records per tile: 14,336 bytes in, 16,384 out
constexpr int kUnpackWordsIn = kUnpackTile * 14 / 4;    // 3,584
__global__ void __launch_bounds__(kUnpackThreads) k_unpack_wire(const uint4 *__restrict__ wire, uint4 *__restrict__ recs, size_t n_recs) {
  __shared__ __align__(16) uint32_t w[kUnpackWordsIn];
  const size_t n_tiles = (n_recs + kUnpackTile - 1) / kUnpackTile;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t r0 = tile * kUnpackTile;
    const int nr = (int)(n_recs - r0 < (size_t)kUnpackTile ? n_recs - r0 : (size_t)kUnpackTile);
    const int nq = (nr * 14 + 15) / 16;                 // (the wire buffer is allocated with 16 bytes of slack)
    const uint4 *src = wire + r0 * 14 / 16;             // a tile starts on a 16-byte boundary (1,024 * 14 = 896 * 16)
    for (int q = threadIdx.x; q < nq; q += kUnpackThreads) reinterpret_cast<uint4 *>(w)[q] = __ldg(src + q);
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += kUnpackThreads) {
      const int sh = (r & 1) << 4;                      // odd records start in the middle of a word
      const uint32_t *p = w + 7 * (r >> 1) + 3 * (r & 1);
      const uint32_t a = p[0], b = p[1], c = p[2], d = p[3];
      uint4 o;
      o.x = __funnelshift_r(a, b, sh); o.y = __funnelshift_r(b, c, sh); o.z = __funnelshift_r(c, d, sh);
      o.w = (d >> sh) & 0xffffu;
      recs[r0 + (size_t)r] = o;                         // bytes 14, 15 (pm_person_site::pad) are zero
    }
    __syncthreads();
  }
}

// VCF input, wire form: three bytes per (record, sample) -- int(PL) capped at 255 of the genotypes (a1,a1), (a1,a2),
// (a2,a2) (FamilyLikelihoodSeq_VCF.cpp:275-279) -- widened to the 16-byte record whose lk[] holds them at the genotype
// indices of the record's two alleles, every other byte 0.  One thread per (record, sample): a warp reads 96 contiguous
// bytes and writes 512.
__global__ void __launch_bounds__(256) k_unpack_pl3(const pm_site_hdr *__restrict__ hdr, const uint8_t *__restrict__ pl3,
                                                    uint4 *__restrict__ recs, size_t n_records, int np) {
  const size_t total = n_records * (size_t)np;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const pm_site_hdr h = hdr[i / (size_t)np];
    const int a1 = h.ref_base, a2 = h.reserved & 0xff;
    const int g[3] = {geno_index(a1, a1), geno_index(a1, a2), geno_index(a2, a2)};
    const uint8_t *p = pl3 + 3 * i;
    uint32_t w[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint32_t v = (uint32_t)__ldg(p + k) << ((g[k] & 3) * 8);
      w[0] |= g[k] < 4 ? v : 0u; w[1] |= (g[k] >= 4 && g[k] < 8) ? v : 0u; w[2] |= g[k] >= 8 ? v : 0u;
    }
    recs[i] = make_uint4(w[0], w[1], w[2], 0u);
  }
}

// ================================================================================================
// microbenchmarks used as roofline denominators
// ================================================================================================
__global__ void k_dfma_peak(double *out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// ================================================================================================
// launchers
// ================================================================================================
cudaError_t launch_sites(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                         const double *d_mono, size_t n_sites, double *d_spill, pm_site_result *d_res, uint16_t *d_status, int *d_err,
                         cudaStream_t stream) {
  if (n_sites == 0) return cudaSuccess;
  if (plan.kind == LaunchPlan::WIDE) return launch_sites_wide(plan, d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, d_res, d_status, d_err, stream);
  const unsigned grid = (unsigned)((n_sites + kNarrowThreads - 1) / kNarrowThreads);
  cudaError_t e = cudaMemsetAsync(d_err + 1, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
#define PM_NARROW(U_, ES_, DN_)                                                                                                                  \
  do {                                                                                                                                         \
    k_sites_narrow<U_, false, ES_, DN_><<<grid, kNarrowThreads, sizeof(NarrowSmem), stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_res, d_status, d_err); \
    k_sites_narrow<U_, true, ES_, DN_><<<grid, kNarrowThreads, sizeof(NarrowSmem), stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_res, d_status, d_err);  \
  } while (0)
  // instances: quartic units only (nuclear families, singletons), units + extended families, extended families only -- the
  // last two with and without the ten-state (--denovo) peel
  if (!plan.es) PM_NARROW(kNarrowMaxUnits, false, true);
  else if (plan.units_per_thread == 0) { if (plan.ten_state) PM_NARROW(0, true, true); else PM_NARROW(0, true, false); }
  else { if (plan.ten_state) PM_NARROW(kNarrowMaxUnits, true, true); else PM_NARROW(kNarrowMaxUnits, true, false); }
#undef PM_NARROW
  return cudaGetLastError();
}

cudaError_t plan_launch(LaunchPlan *plan, int n_person, int n_units, int n_es, int sm_count, const int *force_wide) {
  memset(plan, 0, sizeof *plan);
  plan->n_person = n_person;
  if (n_units <= kNarrowMaxUnits && !force_wide) {
    plan->kind = LaunchPlan::NARROW;
    plan->threads = kNarrowThreads;
    plan->units_per_thread = n_units == 0 ? 0 : kNarrowMaxUnits;
    plan->es = n_es > 0 ? 1 : 0;
    return cudaSuccess;  // (NarrowSmem fits in the default dynamic shared memory limit: no attribute to set)
  }
  return plan_wide(plan, n_person, n_units, n_es, sm_count, force_wide);
}

cudaError_t launch_quick_merge(const uint16_t *d_status_q, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, cudaStream_t stream) {
  if (n_sites == 0) return cudaSuccess;
  k_quick_merge<<<(unsigned)((n_sites + 255) / 256), 256, 0, stream>>>(d_status_q, n_sites, d_res, d_status);
  return cudaGetLastError();
}

cudaError_t launch_compact(const uint16_t *d_status, size_t n_sites, uint32_t *d_emit_sites, uint32_t *d_n_emit, int all,
                           uint32_t *d_tile_scratch, cudaStream_t stream) {
  const size_t n_tiles = (n_sites + 1023) / 1024;
  if (all) k_all_rows<<<(unsigned)std::min<size_t>((n_sites + 255) / 256, 1184), 256, 0, stream>>>(n_sites, d_emit_sites, d_n_emit);
  else if (n_tiles <= 8 || !d_tile_scratch) k_compact<<<1, 1024, 0, stream>>>(d_status, n_sites, d_emit_sites, d_n_emit, 0);
  else {
    k_compact_count<<<(unsigned)n_tiles, 1024, 0, stream>>>(d_status, n_sites, d_tile_scratch);
    k_compact_scan<<<1, 1024, 0, stream>>>(d_tile_scratch, (unsigned)n_tiles, d_n_emit);
    k_compact_scatter<<<(unsigned)n_tiles, 1024, 0, stream>>>(d_status, n_sites, d_tile_scratch, d_emit_sites);
  }
  return cudaGetLastError();
}

cudaError_t launch_unpack_wire(const void *d_wire, uint4 *d_recs, size_t n_recs, int sm_count, cudaStream_t stream) {
  if (n_recs == 0) return cudaSuccess;
  const size_t n_tiles = (n_recs + kUnpackTile - 1) / kUnpackTile;
  const unsigned grid = (unsigned)std::min<size_t>(n_tiles, (size_t)sm_count * 8);  // 8 resident blocks of 256 threads per SM
  k_unpack_wire<<<grid, kUnpackThreads, 0, stream>>>((const uint4 *)d_wire, d_recs, n_recs);
  return cudaGetLastError();
}

cudaError_t launch_unpack_pl3(const pm_site_hdr *d_hdr, const uint8_t *d_pl3, uint4 *d_recs, size_t n_records, int np, int sm_count, cudaStream_t stream) {
  const size_t total = n_records * (size_t)np;
  if (total == 0) return cudaSuccess;
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sm_count * 8);
  k_unpack_pl3<<<grid, 256, 0, stream>>>(d_hdr, d_pl3, d_recs, n_records, np);
  return cudaGetLastError();
}

cudaError_t launch_dfma_peak(double *d_out, int blocks, int threads, int iters, cudaStream_t stream) {
  k_dfma_peak<<<blocks, threads, 0, stream>>>(d_out, iters);
  return cudaGetLastError();
}
cudaError_t launch_copy(const void *src, void *dst, size_t bytes, int sm_count, cudaStream_t stream) {
  k_copy<<<sm_count * 8, 256, 0, stream>>>((const uint4 *)src, (uint4 *)dst, bytes / 16);
  return cudaGetLastError();
}

}  // namespace pm
