// pm_kernels.cu — the site kernels (sm_100a).
//
//  k_sites_narrow : one THREAD per site.  Small pedigrees (<= kNarrowMaxUnits quartic units plus any
//                   number of extended families peeled by Elston–Stewart).  Sites of adjacent lanes are
//                   adjacent in HBM, so a warp streams 32 * n_person * 16 contiguous bytes.
//  k_sites_wide   : one BLOCK per site.  Large pedigrees made of nuclear families and unrelated
//                   founders (hundreds to thousands of units).  The site's n_person*16 bytes are staged
//                   in shared memory with one TMA bulk copy (cp.async.bulk + mbarrier); every thread owns
//                   U units whose quartic coefficients stay in registers for the whole Brent run; one
//                   objective evaluation is 5 FMAs per unit + a (mantissa, exponent) product reduction by
//                   warp shuffles; thread 0 drives the Brent state machine out of shared memory.
//  k_compact      : ordered compaction of the emitted sites (single block scan, deterministic).
//  k_post         : genotype posteriors / GQ / dosage / AB for emitted sites, one thread per (site, family): pm_post.cu.
//
// All arithmetic is FP64; inputs are uint8 phred likelihoods.  See DESIGN.md for the data layout and
// the roofline of each kernel.
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>

#include "pm_device.cuh"
#include "pm_es.cuh"
#include "pm_kernels.h"

namespace pm {

// decisions of main:539-574 once all hypotheses are in; returns true if the de novo refit
// (main:567-573) is needed.  `lk_mono` = MonomorphismLogLikelihood(refBase).
__device__ inline bool site_decide(const DevRun *run, pm_site_result &r, double lk_mono) {
  const int maxidx = r.maxidx;
  if (r.var_post_prob < run->posterior_cutoff) {
    r.flags |= PM_FLAG_NOCALL;
    if (!run->force_call && !run->out_all_sites) { r.status = PM_SITE_NOCALL; return false; }
  }
  if (maxidx == 0) {
    r.freq = 1.0;  // famlk[0].min = 1.0 on every path that reaches the writers (main:544, 561)
  } else {
    int a1, a2;
    hyp_alleles(maxidx, r.reserved /* ref base stashed by the caller */, a1, a2);
    r.allele1 = (uint8_t)a1; r.allele2 = (uint8_t)a2;
    r.freq = r.varfreq[maxidx];
  }
  if (maxidx == 0 && !run->denovo && !run->force_call && !run->out_all_sites) { r.status = PM_SITE_MONO; return false; }
  if (maxidx == 0) {
    if (run->denovo) {
      r.denovo_lr = r.varllk_noprior[0] - lk_mono;
      if (r.denovo_lr <= run->log_min_llr && !run->out_all_sites && !run->force_call) {
        r.status = PM_SITE_DENOVO_LOW_LR;
        return false;
      }
    }
    r.flags |= PM_FLAG_MONO;
    r.status = PM_SITE_EMITTED;
    return false;
  }
  r.status = PM_SITE_EMITTED;
  return run->denovo != 0;
}
__device__ inline void site_finish_refit(const DevRun *run, pm_site_result &r, double lk_poly, double refit_freq) {
  r.refit_llk = lk_poly;
  r.denovo_lr = r.varllk_noprior[r.maxidx] - lk_poly;
  if (run->use_brent) r.freq = refit_freq;  // famlk[0].min is overwritten by the refit's Brent (main:570)
}
__device__ inline void site_store_hyp(const DevRun *run, pm_site_result &r, int h, double maxlogl, double freq, int cls) {
  const double *cl = run->cls_log[cls];
  const double lp = h == 1 ? cl[1] : (h <= 3 ? cl[2] : cl[3]);
  const double ln = h == 1 ? cl[4] : (h <= 3 ? cl[5] : cl[3]);  // main:472,482,492
  double v = lp + maxlogl;
  r.varllk[h] = v;
  r.varllk_noprior[h] = v - ln;
  r.varfreq[h] = freq;
}
__device__ inline uint16_t status_word(const pm_site_result &r) {
  return (uint16_t)(r.status | ((r.maxidx + 1) << 4) | ((r.flags & PM_FLAG_NOCALL) << 8));
}

// ================================================================================================
// narrow kernel: one thread per site
// ================================================================================================
constexpr int kNarrowThreads = 128;

struct NarrowSmem {
  SmemTables t;
  double tden[1000];
  double t10[1000];
};

// NA = the instance for chrX / chrY / MT sites (see k_sites_wide): the autosomal one has none of those rules compiled in.
template <int UMAX, bool NA>
struct NarrowEval {
  const DevRun *run;
  const uint4 *recs;  // this site's records
  const NarrowSmem *sm;
  double B[UMAX][5];
  double C0[9];  // single-nuclear-family mode keeps the nine conditionals
  int g11, g12, g22;
  bool denovo;
  int cls = PM_CHR_AUTO;  // chromosome class of the site; the hypothesis objects' stale `sex` member is 0 (see pm_device.cuh)
  int n_hyp = 0, n_eval = 0;

  __device__ void setup(int a1, int a2, bool dn) {
    g11 = geno_index(a1, a1); g12 = geno_index(a1, a2); g22 = geno_index(a2, a2);
    denovo = dn;
    if (!run->use_brent) {
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
    for (int e = 0; e < run->n_es; e++) {
      const DevFam f = run->fams[run->es_fams[e]];
      double lk = denovo ? es_likelihood_impl<10, NA>(run, f, recs, g11, g12, g22, true, p, sm->t.lut, sm->tden, sm->t10, -1, -1, cls)
                         : es_likelihood_impl<3, NA>(run, f, recs, g11, g12, g22, false, p, sm->t.lut, sm->tden, sm->t10, -1, -1, cls);
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
    if (!run->use_brent) { n_eval++; return loglik_fixed(false); }
    BrentState st;
    brent_begin(st);
    double f;
    do { f = -loglik(st.u); n_eval++; } while (brent_feed(st, f, run->precision));
    *freq = st.min;
    return -st.fmin;
  }
};

template <int UMAX, bool NA>
__global__ void __launch_bounds__(kNarrowThreads) k_sites_narrow(const DevRun *__restrict__ run,
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
  for (int i = threadIdx.x; i < 1000; i += blockDim.x) {
    sm->tden[i] = run->tden[i];
    // transmission[i][j][k] (ES:752-785): a quarter per gamete pair
    int gi = i / 100, gj = (i / 10) % 10, gk = i % 10;
    const int al[10][2] = {{1, 1}, {1, 2}, {1, 3}, {1, 4}, {2, 2}, {2, 3}, {2, 4}, {3, 3}, {3, 4}, {4, 4}};
    double v = 0.0;
    for (int x = 0; x < 2; x++)
      for (int y = 0; y < 2; y++)
        if (geno_index(al[gi][x], al[gj][y]) == gk) v += 0.25;
    sm->t10[i] = v;
  }
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
    NarrowEval<UMAX, NA> ev;
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

  NarrowEval<UMAX, NA> ev;
  ev.run = run; ev.recs = recs; ev.sm = sm; ev.cls = cls;
  r.reserved = (uint16_t)ref;
  // H0 (main:447-462)
  if (!run->denovo) {
    r.varllk[0] = log_1m_prior + lk_mono;
  } else {
    int a1, a2;
    hyp_alleles(0, ref, a1, a2);
    ev.setup(a1, a2, true);
    double l0 = run->use_brent ? ev.loglik(1.0) : ev.loglik_fixed(true);
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
// wide kernel: one block per site, one thread GROUP per Brent chain
//
// One objective evaluation is cheap (5 FMAs per unit) but every Brent step ends in a serial tail —
// product reduction, one log10, the Brent update with its division, two block barriers.  The kernel can
// optimise up to three hypotheses concurrently (G = 3 groups of Tg threads, group g owns chain g, the first
// thread of each group drives its own Brent state); measured on B200, G = 1 with three INDEPENDENT blocks
// per SM is faster (the blocks overlap each other's tails without sharing a barrier), so that is what
// plan_launch picks and G = 3 is only reachable through the PM_WIDE_PLAN tuning hook.  Every thread keeps
// the quartic coefficients of its U units in registers for the whole Brent run.  Under --denovo the hom-ref
// hypothesis H0 is the product of chain 0's p^4 coefficients (the (ref,ref) conditional is the same in H0 and
// H1..H3), so it rides along in the first round.
// ================================================================================================
constexpr int kMaxChains = 3;

#ifdef PM_PHASE_TIMING
#define PM_TICK(slot) do { if (threadIdx.x == 0) { long long now__ = clock64(); ws->phase[slot] += (unsigned long long)(now__ - ws->t_last); ws->t_last = now__; } } while (0)
#else
#define PM_TICK(slot) do { } while (0)
#endif

struct WideShared {
  SmemTables t;
  double log_inv[128];                // table-driven log10 of the product mantissa (driver threads only)
  double log_tab[128];
  pm_site_result r;                   // written by thread 0 only
  BrentState brent[kMaxChains];       // chain c is driven by the first thread of group c
  double p[kMaxChains];               // next evaluation point per chain, < 0 = chain finished
  double warp_m[kMaxChains + 1][32];  // per-warp partial products (slot kMaxChains: the H0 product)
  int warp_e[kMaxChains + 1][32];
  int red_i[4 * 32];                  // per-warp integer partials (depth, samples, mapq, lk sum)
  double bcast[4];
  int ibcast[4];
  int cls;                            // chromosome class of the current site (PM_CHR_*)
  unsigned long long mbar[2];         // one mbarrier per site buffer (TMA bulk copies)
  unsigned int n_hyp, n_eval;         // work counters of the current site
  unsigned int n_eval_g[kMaxChains];  // evaluations per chain driver (summed into n_eval by thread 0)
  unsigned long long phase[8];        // PM_PHASE_TIMING: cycles per phase (thread 0)
  long long t_last;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void tma_issue_site(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// Branch-free running product: acc stays a plain double, rescaled by 2^512 whenever it (or the factor)
// gets small; the power-of-two bookkeeping is an integer.  A factor of exactly 0 makes the product 0,
// i.e. log10 = -inf, which is what the reference computes for a family likelihood that underflowed.
struct FastProd { double m; int e; };
__device__ __forceinline__ void fprod_mul(FastProd &a, double x) {
  const double kBig = 1.3407807929942597e154;   // 2^512
  const double kTiny = 7.458340731200207e-155;  // 2^-512
  const bool xs = x < kTiny;
  x = xs ? x * kBig : x;
  a.e -= xs ? 512 : 0;
  a.m *= x;
  const bool s = a.m < kTiny;
  a.m = s ? a.m * kBig : a.m;
  a.e -= s ? 512 : 0;
}
__device__ __forceinline__ ProdAcc fprod_finish(const FastProd &a) {  // -> mantissa in [1,2) + exponent
  ProdAcc r;
  r.m = a.m; r.e = a.e;
  if (a.m > 0.0) {
    int hi = __double2hiint(r.m);
    int ex = (hi >> 20) & 0x7ff;
    if (ex == 0) { r.m *= 1.3407807929942597e154; r.e -= 512; hi = __double2hiint(r.m); ex = (hi >> 20) & 0x7ff; }
    r.e += ex - 1023;
    r.m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(r.m));
  } else {
    r.m = 0.0;  // log10(0) = -inf downstream
  }
  return r;
}
__device__ __forceinline__ void warp_product(ProdAcc &acc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double m = __shfl_down_sync(0xffffffffu, acc.m, o);
    int e = __shfl_down_sync(0xffffffffu, acc.e, o);
    acc.m *= m; acc.e += e;
  }
}
__device__ __forceinline__ void renorm_nonzero(ProdAcc &a) {
  if (a.m > 0.0) prod_renorm(a);
}

// Out-of-line helpers: their register pressure (unrolled 10-genotype dot products, the Brent update with
// its division, exp10/log10 in the posterior) stays out of the kernel's hot loop allocation.
struct Quartic { double b0, b1, b2, b3, b4; };

// mode = denovo | chr_class << 1 | (single founder's sex) << 3.  NA = the kernel instance for chrX / chrY / MT sites: a
// separate instantiation, so that the autosomal kernel's register allocation does not pay for the rare case (the
// callee's registers count against what the caller can keep live across the call).
template <bool NA>
__device__ __noinline__ Quartic unit_quartic_ol(const uint4 *recs, int first, int nkids, int g11, int g12, int g22, int mode,
                                                const double *lut, const double *mut) {
  double B[5];
  DevUnit u;
  u.first = first; u.nkids = nkids; u.sex = mode >> 3;
  if (!NA) unit_quartic(recs, u, g11, g12, g22, (mode & 1) != 0, lut, mut, B);
  else unit_quartic_nonauto(recs, u, g11, g12, g22, (mode & 1) != 0, (mode >> 1) & 3, 0, lut, mut, B);
  Quartic q;
  q.b0 = B[0]; q.b1 = B[1]; q.b2 = B[2]; q.b3 = B[3]; q.b4 = B[4];
  return q;
}
// Same as unit_quartic for a nuclear family whose kids' mutation-mixed likelihoods
// D[kid][g] = sum_h M[g][h] l_h (CalcDenovoMutLk, NucFam:1553-1562) were built once for the site.
__device__ __noinline__ Quartic unit_quartic_tab_ol(const uint4 *recs, int first, int nkids, const double *kidD, int g11, int g12, int g22,
                                                    const double *lut) {
  double p0 = 1.0, p1 = 1.0, p2 = 1.0, p4 = 1.0, p5 = 1.0, p8 = 1.0;
  for (int k = 0; k < nkids; k++) {
    const double d11 = kidD[k * 10 + g11], d12 = kidD[k * 10 + g12], d22 = kidD[k * 10 + g22];
    p0 *= d11;
    p1 *= 0.5 * (d11 + d12);
    p2 *= d12;
    p4 *= 0.25 * d11 + 0.5 * d12 + 0.25 * d22;
    p5 *= 0.5 * (d12 + d22);
    p8 *= d22;
  }
  const uint4 rf = recs[first], rm = recs[first + 1];
  const double f11 = lut[rec_lk(rf, g11)], f12 = lut[rec_lk(rf, g12)], f22 = lut[rec_lk(rf, g22)];
  const double m11 = lut[rec_lk(rm, g11)], m12 = lut[rec_lk(rm, g12)], m22 = lut[rec_lk(rm, g22)];
  const double C0 = p0 * (f11 * m11), C1 = p1 * (f11 * m12), C2 = p2 * (f11 * m22);
  const double C3 = p1 * (f12 * m11), C4 = p4 * (f12 * m12), C5 = p5 * (f12 * m22);
  const double C6 = p2 * (f22 * m11), C7 = p5 * (f22 * m12), C8 = p8 * (f22 * m22);
  Quartic q;
  q.b4 = C0; q.b3 = 2.0 * (C1 + C3); q.b2 = C2 + 4.0 * C4 + C6; q.b1 = 2.0 * (C5 + C7); q.b0 = C8;
  return q;
}
// One kid's ten D values (the whole row space of the mutation matrix), written to the site's kid table.
__device__ __noinline__ void kid_table_row_ol(const uint4 rk, const double *lut, const double *mut, double *out) {
  double l[10];
#pragma unroll
  for (int g = 0; g < 10; g++) l[g] = lut[rec_lk(rk, g)];
#pragma unroll 2
  for (int x = 0; x < 10; x++) {
    const double *row = mut + x * 10;
    double d = 0.0;
#pragma unroll
    for (int g = 0; g < 10; g++) d += row[g] * l[g];
    out[x] = d;
  }
}
// works on a register copy of the state: one batch of loads, the branchy update without memory traffic, one batch of stores
// Returns the next evaluation point, or -1 when the minimisation is over.
__device__ __noinline__ double brent_feed_ol(BrentState *s, double fu, double tol) {
  BrentState t = *s;
  const bool more = brent_feed(t, fu, tol);
  *s = t;
  return more ? t.u : -1.0;
}
__device__ __noinline__ void var_posterior_ol(pm_site_result *r, int ref, int n) { var_posterior(*r, ref, n); }
__device__ __noinline__ int site_decide_ol(const DevRun *run, pm_site_result *r, double lk_mono) { return site_decide(run, *r, lk_mono) ? 1 : 0; }
// log10(m * 2^e) for a mantissa m in [1,2) (or 0 -> -inf).  m = c (1 + r) with c from a 128-entry table,
// |r| <= 2^-8, log1p(r) by its series to r^7 (truncation < 2^-67); absolute error ~1e-16, far below the
// rounding noise of the reference's own sum of per-family log10 values.  Replaces a ~40-instruction
// dependent libm sequence on the serial tail of every Brent step.
__device__ __noinline__ double log10_ol(const WideShared *ws, double m, int e) {
  if (!(m > 0.0)) return -CUDART_INF;
  const int i = (__double2hiint(m) >> 13) & 127;
  const double r = fma(m, ws->log_inv[i], -1.0);
  // log1p(r) = r + r^2 (c0 + c1 r + ... + c5 r^5), Estrin's scheme: this is on the serial tail of every Brent round
  const double r2 = r * r;
  const double a = fma(1.0 / 3.0, r, -0.5), b = fma(1.0 / 5.0, r, -0.25), c = fma(1.0 / 7.0, r, -1.0 / 6.0);
  const double r4 = r2 * r2;
  const double q = fma(c, r4, fma(b, r2, a));
  const double l1p = fma(r2, q, r);
  return fma((double)e, kLog10_2, fma(l1p, 0.43429448190325182765, ws->log_tab[i]));
}

// ES instances: extended families take part in every evaluation through a thread-serial Elston-Stewart peel, family e
// on thread e of the group (FLSeq.cpp:222-240 sums log10 over all families; here: one more factor of the product).
// A free function on purpose, and only ever named under `if constexpr (ES)`: a member function would let `this`
// escape, the evaluator's pointers would lose their address space and the hot loop's LDS turn into generic LD
// (measured: -15 % on the autosomal instance even though it never calls this).
__device__ __noinline__ double es_factor(const DevRun *run, const uint4 *recs, const double *lut, int cls, int e, int g11, int g12, int g22,
                                         bool denovo, double p) {
  const DevFam f = run->fams[run->es_fams[e]];
  return denovo ? es_likelihood<10>(run, f, recs, g11, g12, g22, true, p, lut, run->tden, run->t10, -1, -1, cls)
                : es_likelihood<3>(run, f, recs, g11, g12, g22, false, p, lut, run->tden, run->t10, -1, -1, cls);
}

template <int U, bool NA, bool ES>
struct WideEval {
  const DevRun *run;
  const uint4 *recs;  // site records in shared memory
  WideShared *ws;
  const double *kidD;  // per-site kid table in shared memory (nullptr = build D on the fly)
  int G, Tg, grp, t;  // groups, threads per group, my group, my index inside the group
  int eg11, eg12, eg22;  // ES: genotype indices of the current hypothesis (the peel has no per-hypothesis set-up)
  double B[U][5];     // per unit, scaled by an exact power of two so that the largest coefficient is in [1,2)
  int K;              // sum over my units of the exponents taken out: prod_u L_u = 2^K * prod_u L'_u

  // Optimises nc <= G hypotheses concurrently (chain c by group c).  On return (after a barrier)
  // ws->brent[c] holds min/fmin of chain c.  with_h0: also reduce prod_u B[u][4] of chain 0 in the first
  // round; its log10 lands in ws->bcast[1].
  __device__ __forceinline__ void optimize(int nc, const int *a1, const int *a2, bool denovo, bool with_h0) {
    const int lane = threadIdx.x & 31;
    const int wg = t >> 5, nwg = Tg >> 5;  // warp index inside the group, warps per group
    const bool mine = grp < nc;
    K = 0;
    if (mine) {
      const int x = a1[grp], y = a2[grp];
      const int g11 = geno_index(x, x), g12 = geno_index(x, y), g22 = geno_index(y, y);
      if constexpr (ES) { eg11 = g11; eg12 = g12; eg22 = g22; }
      const int mode = (denovo ? 1 : 0) | (NA ? (ws->cls << 1) : 0);
#pragma unroll
      for (int k = 0; k < U; k++) {
        const int u = t + k * Tg;
        if (u < run->n_units) {
          const DevUnit du = run->units[u];
          const Quartic q = (denovo && kidD && du.nkids > 0)
                                ? unit_quartic_tab_ol(recs, du.first, du.nkids, kidD + (size_t)du.kid0 * 10, g11, g12, g22, ws->t.lut)
                                : unit_quartic_ol<NA>(recs, du.first, du.nkids, g11, g12, g22, NA ? (mode | (du.sex << 3)) : mode, ws->t.lut, ws->t.mut);
          // With the largest coefficient in [1,2) and p in [1e-4, 0.9999], L'(p) >= min(p,q)^4 >= 2^-54, so the
          // product over this thread's U <= 8 units cannot underflow: no exponent handling inside a round.
          double mx = fmax(fmax(fmax(q.b0, q.b1), fmax(q.b2, q.b3)), q.b4);
          double b0 = q.b0, b1 = q.b1, b2 = q.b2, b3 = q.b3, b4 = q.b4;
          if (mx > 0.0 && mx < 2.2250738585072014e-308) {  // subnormal maximum: lift it first
            const double big = 1.3407807929942597e154;     // 2^512
            b0 *= big; b1 *= big; b2 *= big; b3 *= big; b4 *= big; mx *= big;
            K -= 512;
          }
          if (mx > 0.0) {
            const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
            const double sc = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023-ex), exact
            b0 *= sc; b1 *= sc; b2 *= sc; b3 *= sc; b4 *= sc;
            K += ex - 1023;
          }
          B[k][0] = b0; B[k][1] = b1; B[k][2] = b2; B[k][3] = b3; B[k][4] = b4;
        } else {  // (p+q)^4 = 1: a neutral unit
          B[k][4] = 1.0; B[k][3] = 4.0; B[k][2] = 6.0; B[k][1] = 4.0; B[k][0] = 1.0;
        }
      }
    }
    if (t == 0) {
      if (mine) { brent_begin(ws->brent[grp]); ws->p[grp] = ws->brent[grp].u; }
      else ws->p[grp] = -1.0;
    }
    if (threadIdx.x == 0) {
      for (int c = G; c < kMaxChains; c++) ws->p[c] = -1.0;
      ws->n_hyp += nc;
    }
    __syncthreads();
    PM_TICK(2);
    bool first = true;
    unsigned n_eval_mine = 0;  // rounds driven by this thread (the first of its group)
    for (;;) {
      const double p0 = ws->p[0], p1 = ws->p[1], p2 = ws->p[2];
      if (p0 < 0.0 && p1 < 0.0 && p2 < 0.0) break;
      const double p = grp == 0 ? p0 : (grp == 1 ? p1 : p2);
      const bool live = p >= 0.0;
      const bool h0 = first && with_h0 && grp == 0;
      if (live) {
        const Monomials m = monomials(p);
        FastProd fa;
        fa.e = K;
        {  // four independent partial products keep the dependency chain short
          double v0 = 1.0, v1 = 1.0, v2 = 1.0, v3 = 1.0;
#pragma unroll
          for (int k = 0; k < U; k += 4) {
            v0 *= quartic_eval(B[k], m);
            if (k + 1 < U) v1 *= quartic_eval(B[k + 1], m);
            if (k + 2 < U) v2 *= quartic_eval(B[k + 2], m);
            if (k + 3 < U) v3 *= quartic_eval(B[k + 3], m);
          }
          fa.m = (v0 * v1) * (v2 * v3);
        }
        if constexpr (ES)
          for (int e = t; e < run->n_es; e += Tg)
            fprod_mul(fa, es_factor(run, recs, ws->t.lut, NA ? ws->cls : PM_CHR_AUTO, e, eg11, eg12, eg22, denovo, p));
        ProdAcc acc = fprod_finish(fa);
        warp_product(acc);
        if (lane == 0) { renorm_nonzero(acc); ws->warp_m[grp][wg] = acc.m; ws->warp_e[grp][wg] = acc.e; }
      }
      if (h0) {
        FastProd fa;
        fa.m = 1.0; fa.e = K;   // B[k][4] carries the same power-of-two scaling as the unit
#pragma unroll
        for (int k = 0; k < U; k++) fprod_mul(fa, B[k][4]);
        if constexpr (ES)  // a founder's prior at p = 1 is (1, 0, 0) whatever the second allele: H1's alleles give H0's value
          for (int e = t; e < run->n_es; e += Tg)
            fprod_mul(fa, es_factor(run, recs, ws->t.lut, NA ? ws->cls : PM_CHR_AUTO, e, eg11, eg12, eg22, true, 1.0));
        ProdAcc acc = fprod_finish(fa);
        warp_product(acc);
        if (lane == 0) { renorm_nonzero(acc); ws->warp_m[kMaxChains][wg] = acc.m; ws->warp_e[kMaxChains][wg] = acc.e; }
      }
      __syncthreads();
      PM_TICK(3);
      // serial tails: the first thread of every live group, in parallel
      if (t == 0 && live) {
        ProdAcc a;
        {  // the warps' mantissas are in [1,2): two running products halve the dependent chain
          double m0 = 1.0, m1 = 1.0;
          int e = 0;
          for (int w = 0; w < nwg; w += 2) {
            m0 *= ws->warp_m[grp][w]; e += ws->warp_e[grp][w];
            if (w + 1 < nwg) { m1 *= ws->warp_m[grp][w + 1]; e += ws->warp_e[grp][w + 1]; }
          }
          a.m = m0 * m1; a.e = e;
        }
        renorm_nonzero(a);  // log10_ol wants the mantissa back in [1,2)
        const double ll = log10_ol(ws, a.m, a.e);
        ws->p[grp] = brent_feed_ol(&ws->brent[grp], -ll, run->precision);
        n_eval_mine++;
      }
      if (h0 && t == (Tg > 32 ? 32 : 0)) {
        ProdAcc a;
        a.m = 1.0; a.e = 0;
        for (int w = 0; w < nwg; w++) prod_merge(a, ws->warp_m[kMaxChains][w], ws->warp_e[kMaxChains][w]);
        renorm_nonzero(a);
        ws->bcast[1] = log10_ol(ws, a.m, a.e);
      }
      first = false;
      __syncthreads();
      PM_TICK(4);
    }
    if (t == 0 && mine) ws->n_eval_g[grp] += n_eval_mine;  // read by thread 0 after the site's last barrier
  }
};

// NA = false: the autosomal instance; sites on chrX / chrY / MT are left untouched and flagged in err[1].
// NA = true: launched right behind it, returns at once unless err[1] is set, then does only those sites.
// ES = the pedigree also has extended families (evaluated by WideEval::es_factor); again separate instances.
template <int U, int MAXT, bool NA, bool ES>
__global__ void __launch_bounds__(MAXT, 1) k_sites_wide(const DevRun *__restrict__ run, const pm_site_hdr *__restrict__ hdr,
                                                        const uint4 *__restrict__ recs_all, const double *__restrict__ mono_all,
                                                        size_t n_sites, int groups, int nbuf, int kid_table,
                                                        pm_site_result *__restrict__ res, uint16_t *__restrict__ status,
                                                        int *__restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WideShared *ws = reinterpret_cast<WideShared *>(smem_raw);
  if (NA ? ((err[1] == 0 && run->site_filter != 2) || run->site_filter == 1) : run->site_filter == 2) return;
  const int np = run->n_person;
  const size_t site_bytes = (((size_t)np * 16 + 127) / 128) * 128;
  unsigned char *site_base = smem_raw + ((sizeof(WideShared) + 127) / 128) * 128;
  double *kid_tab = kid_table ? reinterpret_cast<double *>(site_base + (size_t)nbuf * site_bytes) : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  load_tables(run, &ws->t);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { ws->log_inv[i] = run->log_inv[i]; ws->log_tab[i] = run->log_tab[i]; }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 8; k++) ws->phase[k] = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ws->mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ws->mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // nbuf = 2: the next site's n_person*16 bytes are fetched by TMA while this one is being computed
  uint32_t phase[2] = {0, 0};
  int cur = 0;
  if (threadIdx.x == 0 && blockIdx.x < n_sites)
    tma_issue_site(site_base, recs_all + (size_t)blockIdx.x * np, (uint32_t)np * 16u, &ws->mbar[0]);
  WideEval<U, NA, ES> ev;
  ev.run = run; ev.ws = ws; ev.kidD = kid_tab;
  ev.G = groups; ev.Tg = blockDim.x / groups; ev.grp = threadIdx.x / ev.Tg; ev.t = threadIdx.x % ev.Tg;
  const int G = groups;

  for (size_t s = blockIdx.x; s < n_sites; s += gridDim.x) {
    // ---- the site: n_person * 16 contiguous bytes, one TMA bulk copy (issued one iteration ahead) ----
    uint4 *site = reinterpret_cast<uint4 *>(site_base + (size_t)cur * site_bytes);
    ev.recs = site;
    if (nbuf == 2) {
      const size_t nxt = s + gridDim.x;
      if (threadIdx.x == 0 && nxt < n_sites)
        tma_issue_site(site_base + (size_t)(cur ^ 1) * site_bytes, recs_all + nxt * (size_t)np, (uint32_t)np * 16u, &ws->mbar[cur ^ 1]);
    }
#ifdef PM_PHASE_TIMING
    if (threadIdx.x == 0) ws->t_last = clock64();
#endif
    mbar_wait(&ws->mbar[cur], phase[cur]);
    phase[cur] ^= 1;
    PM_TICK(0);
    const pm_site_hdr h = hdr[s];
    const int ref = h.ref_base;
    bool skip = false;
    const int cls = h.chr_class;
    const bool bad_cls = cls > PM_CHR_MT;
    if (NA && threadIdx.x == 0) ws->cls = cls;  // read by the set-up loops after the next barrier
    const bool bad = ref < 1 || ref > 4 || bad_cls;
    if (NA ? (bad || cls == PM_CHR_AUTO) : (!bad && cls != PM_CHR_AUTO)) {  // the other instance's site
      if (!NA && threadIdx.x == 0) atomicExch(err + 1, 1);
      skip = true;
    } else if (bad) {
      if (threadIdx.x == 0) {
        pm_site_result &r = ws->r;
        memset(&r, 0, sizeof r);
        r.site = (uint32_t)s; r.maxidx = -1; r.status = PM_SITE_BAD_REF;
        if (ref >= 1 && ref <= 4) atomicExch(err, PM_EUNSUPPORTED);
        res[s] = r; status[s] = status_word(r);
      }
      skip = true;
    }
    if (!skip && run->vcf_mode) {  // one record of a VCF: mono is given, one Brent run for (REF, ALT)
      const int a1 = ref, a2 = h.reserved & 0xff;
      if (threadIdx.x == 0) {
        memset(&ws->r, 0, sizeof ws->r);
        ws->r.site = (uint32_t)s;
        ws->n_hyp = 0; ws->n_eval = 0;
        for (int c = 0; c < kMaxChains; c++) ws->n_eval_g[c] = 0;
      }
      ev.optimize(1, &a1, &a2, false, false);
      if (threadIdx.x == 0) {
        vcf_record_result(run, ws->r, a1, a2, (h.reserved & 0x100) != 0, mono_all[s], -ws->brent[0].fmin, ws->brent[0].min);
        res[s] = ws->r;
        status[s] = status_word(ws->r);
      }
      skip = true;
    }
    if (!skip) {
      // ---- CalcReadStats / MonomorphismLogLikelihood: integer block reductions ----
      const int grr = geno_index(ref, ref);
      int dsum = 0, nsamp = 0, mq = 0, lksum = 0;
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        uint4 rec = site[i];
        int d = rec_depth(rec);
        dsum += d; nsamp += d > 0; mq += rec_mapq(rec); lksum += (int)rec_lk(rec, grr);
      }
      dsum = __reduce_add_sync(0xffffffffu, dsum); nsamp = __reduce_add_sync(0xffffffffu, nsamp);
      mq = __reduce_add_sync(0xffffffffu, mq); lksum = __reduce_add_sync(0xffffffffu, lksum);
      if (lane == 0) { ws->red_i[warp] = dsum; ws->red_i[32 + warp] = nsamp; ws->red_i[64 + warp] = mq; ws->red_i[96 + warp] = lksum; }
      __syncthreads();
      if (threadIdx.x == 0) {
        int D = 0, NS = 0, MQ = 0, LK = 0;
        for (int w = 0; w < nwarp; w++) { D += ws->red_i[w]; NS += ws->red_i[32 + w]; MQ += ws->red_i[64 + w]; LK += ws->red_i[96 + w]; }
        pm_site_result &r = ws->r;
        memset(&r, 0, sizeof r);
        r.site = (uint32_t)s; r.maxidx = -1;
        r.total_depth = D; r.num_samp = NS;
        if (NS > 0) { r.avg_map_qual = (double)MQ / (double)NS; r.perc_samp = (double)NS / (double)np; }
        if (D < run->min_total_depth) r.status = PM_SITE_MIN_DEPTH;
        else if (run->max_total_depth > 0 && D > run->max_total_depth) r.status = PM_SITE_MAX_DEPTH;
        else if (r.perc_samp * 100 < run->min_ps) r.status = PM_SITE_MIN_PS;
        else if (r.avg_map_qual < run->min_map_quality) r.status = PM_SITE_MIN_MAPQ;
        // sum_i -lk_i/10 with the integer sum taken first (exact), one division
        ws->bcast[0] = -(double)LK / 10.0;
        ws->ibcast[0] = r.status;
        r.reserved = (uint16_t)ref;
        ws->n_hyp = 0; ws->n_eval = 0;
        for (int c = 0; c < kMaxChains; c++) ws->n_eval_g[c] = 0;
        if (r.status != 0) { res[s] = r; status[s] = status_word(r); }
      }
      __syncthreads();
      PM_TICK(1);
      skip = ws->ibcast[0] != 0;
    }
    if (!skip) {
      const double lk_mono = ws->bcast[0];
      const bool dn = run->denovo != 0;
      if (kid_tab) {
        // once per site: every kid's ten mutation-mixed likelihoods, shared by all hypotheses of the site
        for (int u = threadIdx.x; u < run->n_units; u += blockDim.x) {
          const DevUnit du = run->units[u];
          for (int k = 0; k < du.nkids; k++) kid_table_row_ol(site[du.first + 2 + k], ws->t.lut, ws->t.mut, kid_tab + (size_t)(du.kid0 + k) * 10);
        }
        __syncthreads();
      }
      // ---- H1..H3 (+ H0 under --denovo), then H4..H6 if the posterior is not decisive ----
      for (int base = 1; base <= 4; base += 3) {
        int a1[3], a2[3];
        for (int c = 0; c < 3; c++) hyp_alleles(base + c, ref, a1[c], a2[c]);
        for (int c0 = 0; c0 < 3; c0 += G) {
          const int nc = (3 - c0) < G ? (3 - c0) : G;
          const bool with_h0 = dn && base == 1 && c0 == 0;
          ev.optimize(nc, a1 + c0, a2 + c0, dn, with_h0);
          if (threadIdx.x == 0) {
            for (int c = 0; c < nc; c++) site_store_hyp(run, ws->r, base + c0 + c, -ws->brent[c].fmin, ws->brent[c].min, cls);
            if (with_h0) { ws->r.varllk[0] = run->cls_log[cls][0] + ws->bcast[1]; ws->n_hyp += 1; ws->n_eval += 1; }
          }
        }
        if (threadIdx.x == 0) {
          if (base == 1) {
            if (!dn) ws->r.varllk[0] = run->cls_log[cls][0] + lk_mono;
            ws->r.varllk_noprior[0] = ws->r.varllk[0] - run->cls_log[cls][0];
            ws->r.varfreq[0] = 1.0;
            var_posterior_ol(&ws->r, ref, 4);
            ws->ibcast[1] = ws->r.var_post_prob < 0.99;  // main:499
          } else {
            var_posterior_ol(&ws->r, ref, 7);
          }
        }
        __syncthreads();
        if (!ws->ibcast[1]) break;
      }
      if (threadIdx.x == 0) ws->ibcast[2] = site_decide_ol(run, &ws->r, lk_mono);
      __syncthreads();
      if (ws->ibcast[2]) {  // de novo refit without mutation (main:567-573)
        const int a1 = ws->r.allele1, a2 = ws->r.allele2;
        ev.optimize(1, &a1, &a2, false, false);
        if (threadIdx.x == 0) site_finish_refit(run, ws->r, -ws->brent[0].fmin, ws->brent[0].min);
      }
      if (threadIdx.x == 0) {
        pm_site_result &r = ws->r;
        if (r.status == PM_SITE_EMITTED && run->denovo && r.denovo_lr < run->denovo_min_llr) { r.flags |= PM_FLAG_ROW_DROPPED; r.status = PM_SITE_DENOVO_DROPPED; }
        r.reserved = 0;
        status[s] = status_word(r);
        atomicAdd(&run->counters[0], (unsigned long long)ws->n_hyp);
        atomicAdd(&run->counters[1], (unsigned long long)(ws->n_eval + ws->n_eval_g[0] + ws->n_eval_g[1] + ws->n_eval_g[2]));
        atomicAdd(&run->counters[2], 1ull);
        atomicAdd(&run->counters[3], (unsigned long long)(r.status == PM_SITE_EMITTED));
#ifdef PM_PHASE_TIMING
        PM_TICK(5);
        for (int k = 0; k < 8; k++) { atomicAdd(&run->counters[8 + k], ws->phase[k]); ws->phase[k] = 0; }
#endif
      }
    }
    __syncthreads();  // ws->r is final
    if (!skip && warp == 0) {  // the 256-byte result leaves as one coalesced 32 x 8-byte store
      static_assert(sizeof(pm_site_result) == 256, "pm_site_result must be 256 bytes");
      reinterpret_cast<unsigned long long *>(&res[s])[lane] = reinterpret_cast<const unsigned long long *>(&ws->r)[lane];
    }
    __syncthreads();  // the site buffer and ws->r are reused by the next iteration
    if (nbuf == 2) cur ^= 1;
    else if (threadIdx.x == 0 && s + gridDim.x < n_sites)
      tma_issue_site(site_base, recs_all + (s + gridDim.x) * (size_t)np, (uint32_t)np * 16u, &ws->mbar[0]);
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
static size_t wide_smem_bytes(int n_person, int nbuf, int n_kids_table) {
  return ((sizeof(WideShared) + 127) / 128) * 128 + (size_t)nbuf * ((((size_t)n_person * 16 + 127) / 128) * 128) + (size_t)n_kids_table * 80 + 16;
}

// (U, MAXT) instantiations of the wide kernel.  U = units per thread; MAXT = largest block the
// instantiation is launched with (sets the register budget: 65536 / MAXT).
#define PM_WIDE_DISPATCH(plan_, CALL)                                   \
  do {                                                                  \
    switch ((plan_).units_per_thread) {                                 \
      case 1: CALL(1, 1024); break;                                     \
      case 2: CALL(2, 768); break;                                      \
      case 4: if ((plan_).threads > 512 || (plan_).low_regs) { CALL(4, 768); } else { CALL(4, 512); } break; \
      case 8: if ((plan_).threads > 384 || (plan_).low_regs) { CALL(8, 512); } else { CALL(8, 384); } break; \
      default: CALL(16, 128); break;                                    \
    }                                                                   \
  } while (0)

cudaError_t launch_sites(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                         const double *d_mono, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, int *d_err,
                         cudaStream_t stream) {
  if (n_sites == 0) return cudaSuccess;
  if (plan.kind == LaunchPlan::NARROW) {
    const unsigned grid = (unsigned)((n_sites + kNarrowThreads - 1) / kNarrowThreads);
    cudaError_t e = cudaMemsetAsync(d_err + 1, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    k_sites_narrow<kNarrowMaxUnits, false><<<grid, kNarrowThreads, sizeof(NarrowSmem), stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_res, d_status, d_err);
    k_sites_narrow<kNarrowMaxUnits, true><<<grid, kNarrowThreads, sizeof(NarrowSmem), stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_res, d_status, d_err);
  } else {
    const size_t smem = wide_smem_bytes(plan.n_person, plan.site_buffers, plan.kid_table);
    const unsigned grid = (unsigned)(n_sites < (size_t)plan.grid ? n_sites : (size_t)plan.grid);
    cudaError_t e = cudaMemsetAsync(d_err + 1, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
#define PM_WIDE_ARGS d_run, d_hdr, d_recs, d_mono, n_sites, plan.chains, plan.site_buffers, plan.kid_table, d_res, d_status, d_err
#define PM_WIDE(U_, MT_)                                                                       \
  if (plan.es) {                                                                               \
    k_sites_wide<U_, MT_, false, true><<<grid, plan.threads, smem, stream>>>(PM_WIDE_ARGS);    \
    k_sites_wide<U_, MT_, true, true><<<grid, plan.threads, smem, stream>>>(PM_WIDE_ARGS);     \
  } else {                                                                                     \
    k_sites_wide<U_, MT_, false, false><<<grid, plan.threads, smem, stream>>>(PM_WIDE_ARGS);   \
    k_sites_wide<U_, MT_, true, false><<<grid, plan.threads, smem, stream>>>(PM_WIDE_ARGS);    \
  }
    PM_WIDE_DISPATCH(plan, PM_WIDE);
#undef PM_WIDE
#undef PM_WIDE_ARGS
  }
  return cudaGetLastError();
}

cudaError_t plan_launch(LaunchPlan *plan, int n_person, int n_units, int n_es, int n_kids_denovo, int sm_count) {
  plan->n_person = n_person;
  plan->chains = 1;
  plan->kid_table = 0;
  plan->site_buffers = 1;
  plan->low_regs = 0;
  plan->es = 0;
  if (n_units <= kNarrowMaxUnits) {
    plan->kind = LaunchPlan::NARROW;
    plan->threads = kNarrowThreads;
    plan->units_per_thread = kNarrowMaxUnits;
    plan->grid = 0;
    plan->blocks_per_sm = 0;
    cudaError_t e = cudaFuncSetAttribute(k_sites_narrow<kNarrowMaxUnits, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NarrowSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_sites_narrow<kNarrowMaxUnits, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NarrowSmem));
  }
  plan->kind = LaunchPlan::WIDE;
  plan->es = n_es > 0 ? 1 : 0;  // extended families ride along as thread-serial peels (WideEval::es_factor)
  // G groups (one Brent chain each) x Tg threads x U units per thread, Tg * U >= n_units.
  // Measured on B200 (1,000 trios, --denovo): one chain per block with three independent blocks per SM
  // (6.7 M sites/s) beats three chains in one 384-thread block (5.6 M): independent blocks overlap each
  // other's serial Brent tails without sharing a barrier.  So G = 1 and as many resident blocks as fit.
  int U, Tg, G = 1;
  auto up32 = [](int x) { return ((x + 31) / 32) * 32; };
  if (n_units <= 32) { U = 1; Tg = 32; }
  else if (n_units <= 64) { U = 2; Tg = 32; }
  else if (n_units <= 128) { U = 4; Tg = 32; }
  else if (n_units <= 4096) { U = 8; Tg = up32((n_units + 7) / 8); }
  else return cudaErrorNotSupported;
  // tuning hook: PM_WIDE_PLAN="threads_per_group,units_per_thread,groups" overrides the choice
  if (const char *env = getenv("PM_WIDE_PLAN")) {
    int t = 0, u = 0, g = 0;
    if (sscanf(env, "%d,%d,%d", &t, &u, &g) == 3 && t >= 32 && t % 32 == 0 && (g == 1 || g == 3) && (long)t * u >= n_units &&
        (u == 1 || u == 2 || u == 4 || u == 8 || u == 16) && t * g <= (u == 1 ? 1024 : (u == 16 ? 128 : (u == 8 ? 512 : 768)))) {
      Tg = t; U = u; G = g;
    }
  }
  plan->threads = Tg * G;
  plan->units_per_thread = U;
  plan->chains = G;
  // shared-memory budget: prefer the kid table (saves set-up arithmetic on every hypothesis), then the second
  // site buffer (hides the TMA latency)
  const size_t budget = 220 * 1024;
  plan->kid_table = (n_kids_denovo > 0 && wide_smem_bytes(n_person, 1, n_kids_denovo) <= budget) ? n_kids_denovo : 0;
  // a second site buffer (TMA prefetch of the next site) only pays when it does not cost a resident block
  plan->site_buffers = wide_smem_bytes(n_person, 2, plan->kid_table) <= 24 * 1024 ? 2 : 1;
  if (getenv("PM_WIDE_TWO_BUF") && wide_smem_bytes(n_person, 2, plan->kid_table) <= budget) plan->site_buffers = 2;
  if (!getenv("PM_WIDE_KIDTAB")) plan->kid_table = 0;  // measured: the 10-row table costs as much as it saves
  plan->low_regs = getenv("PM_WIDE_LOWREGS") ? 1 : 0;
  if (getenv("PM_WIDE_ONE_BUF")) plan->site_buffers = 1;
  const size_t smem = wide_smem_bytes(n_person, plan->site_buffers, plan->kid_table);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  cudaError_t e = cudaSuccess;
  int per_sm = 1;
  const int T = plan->threads;
#define PM_ATTR(U_, MT_)                                                                                              \
  if (plan->es) {                                                                                                        \
    e = cudaFuncSetAttribute(k_sites_wide<U_, MT_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sites_wide<U_, MT_, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sites_wide<U_, MT_, false, true>, T, smem); \
  } else {                                                                                                               \
    e = cudaFuncSetAttribute(k_sites_wide<U_, MT_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sites_wide<U_, MT_, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sites_wide<U_, MT_, false, false>, T, smem); \
  }
  PM_WIDE_DISPATCH(*plan, PM_ATTR);
#undef PM_ATTR
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  plan->grid = sm_count * per_sm;  // persistent: a multiple of the SM count
  plan->blocks_per_sm = per_sm;
  return cudaSuccess;
}

cudaError_t launch_quick_merge(const uint16_t *d_status_q, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, cudaStream_t stream) {
  if (n_sites == 0) return cudaSuccess;
  k_quick_merge<<<(unsigned)((n_sites + 255) / 256), 256, 0, stream>>>(d_status_q, n_sites, d_res, d_status);
  return cudaGetLastError();
}

cudaError_t launch_compact(const uint16_t *d_status, size_t n_sites, uint32_t *d_emit_sites, uint32_t *d_n_emit, int all,
                           cudaStream_t stream) {
  k_compact<<<1, 1024, 0, stream>>>(d_status, n_sites, d_emit_sites, d_n_emit, all);
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
