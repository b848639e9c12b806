// pm_site_logic.cuh — per-site decisions shared by the narrow and the wide site kernels (main.cpp:439-594).
#pragma once
#include "pm_device.cuh"

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

}  // namespace pm
