// pm_post.cu — genotype posteriors / GQ / dosage / AB for the emitted sites (sm_100a).
//
// Built with -fmad=false: a genotype call is an argmax over posteriors that can tie EXACTLY (a haploid male on chrX
// with PL 112,0,112, two untouched transversion alleles ...).  The reference resolves such ties by the rounding of its
// own unfused multiply-then-add sequence (x86-64 without FMA contraction), so this kernel keeps the same operation
// order and lets no multiply-add pair be contracted.  <= 1 % of the sites reach it; it is not on the roofline.
#include <cstdio>
#include <cstdlib>

#include "pm_device.cuh"
#include "pm_es.cuh"
#include "pm_kernels.h"

namespace pm {

// ================================================================================================
// posteriors for emitted sites: one thread per (row, family)
// ================================================================================================
struct PostSmem {
  SmemTables t;
};

__device__ inline void store_person3(pm_person_result &o, double p0, double p1, double p2, int best) {
  o.post[0] = p0; o.post[1] = p1; o.post[2] = p2;
  for (int g = 3; g < 10; g++) o.post[g] = 0.0;
  o.dosage = p1 + p2 * 2;
  o.best = best;
  o.gq = gq_of(best == 0 ? p0 : (best == 1 ? p1 : p2));
  o.ten_state = 0;
  o.reserved[0] = o.reserved[1] = 0;
}

// Where a person's result goes: the 96-byte pm_person_result row, or (CALLS: what the --in_vcf writer prints from,
// FamilyLikelihoodSeq_VCF.cpp:499-517) two bytes, best | gq << 8 -- same arithmetic up to the stored values either way.
template <bool CALLS>
struct PersonSink {
  pm_person_result *out;
  uint16_t *calls;
  __device__ __forceinline__ void put3(int i, double p0, double p1, double p2, int best) const {
    if (CALLS) calls[i] = (uint16_t)(best | ((unsigned)gq_of(best == 0 ? p0 : (best == 1 ? p1 : p2)) << 8));
    else store_person3(out[i], p0, p1, p2, best);
  }
  __device__ __forceinline__ void put10(int i, const double *post, int best) const {
    if (CALLS) { calls[i] = (uint16_t)(best | ((unsigned)gq_of(post[best]) << 8)); return; }
    pm_person_result &o = out[i];
    for (int g = 0; g < 10; g++) o.post[g] = post[g];
    o.dosage = 0.0; o.best = best; o.gq = gq_of(post[best]); o.ten_state = 1;
    o.reserved[0] = o.reserved[1] = 0;
  }
  __device__ __forceinline__ void zero(int i) const {
    if (CALLS) calls[i] = 0;
    else memset(&out[i], 0, sizeof(pm_person_result));
  }
};

// likelihoodKidGenotype on chrX / chrY / MT (NucFam:1334-1443) for configurations 1, 2, 6, 7 (0 and 8 have no
// special case there, 3..5 are zero); `sex` is the kid's own.
__device__ inline void kid_cfg_nonauto(int cls, int sex, int cfg, double l11, double l12, double l22, double &lk, double &x11,
                                       double &x12, double &x22) {
  const bool male = sex == 1;
  x11 = x12 = x22 = 0.0;
  if (cls == PM_CHR_X) {
    switch (cfg) {
      case 1: if (male) { lk = 0.5 * (l11 + l22); x11 = 0.5 * l11; x22 = 0.5 * l22; } else { lk = 0.5 * (l11 + l12); x11 = 0.5 * l11; x12 = 0.5 * l12; } break;
      case 2: if (male) { lk = l22; x22 = l22; } else { lk = l12; x12 = l12; } break;
      case 6: if (male) { lk = l11; x11 = l11; } else { lk = l12; x12 = l12; } break;
      default: if (male) { lk = 0.5 * (l11 + l22); x11 = 0.5 * l11; x22 = 0.5 * l22; } else { lk = 0.5 * (l12 + l22); x12 = 0.5 * l12; x22 = 0.5 * l22; } break;
    }
  } else if (cls == PM_CHR_Y) {
    if (!male) { lk = 1.0; return; }
    if (cfg == 1 || cfg == 2) { lk = l11; x11 = l11; } else { lk = l22; x22 = l22; }
  } else {
    switch (cfg) {
      case 1: case 7: lk = 0.5 * (l11 + l22); x11 = 0.5 * l11; x22 = 0.5 * l22; break;
      case 2: lk = l22; x22 = l22; break;
      default: lk = l11; x11 = l11; break;
    }
  }
}

// likelihoodKidGenotype (NucFam:798-835, 1334-1443) of one kid under parental configuration cfg: lk = the kid's likelihood,
// x11 / x12 / x22 its split by the kid's own genotype.  na: chrX / chrY / MT rules (sex = the kid's).
__device__ __forceinline__ void kid_cfg(int cfg, bool na, int cls, int sex, double l11, double l12, double l22, double &lk, double &x11,
                                        double &x12, double &x22) {
  if (na && cfg != 0 && cfg != 8) {
    if (cfg >= 3 && cfg <= 5) { lk = 0.0; x11 = x12 = x22 = 0.0; }
    else kid_cfg_nonauto(cls, sex, cfg, l11, l12, l22, lk, x11, x12, x22);
    return;
  }
  switch (cfg) {
    case 0: lk = l11; x11 = l11; x12 = 0; x22 = 0; break;
    case 1: case 3: lk = 0.5 * (l11 + l12); x11 = l11 * 0.5; x12 = l12 * 0.5; x22 = 0; break;
    case 2: case 6: lk = l12; x11 = 0; x12 = l12; x22 = 0; break;
    case 4: lk = 0.25 * l11 + 0.5 * l12 + 0.25 * l22; x11 = l11 * 0.25; x12 = l12 * 0.5; x22 = l22 * 0.25; break;
    case 5: case 7: lk = 0.5 * (l12 + l22); x11 = 0; x12 = l12 * 0.5; x22 = l22 * 0.5; break;
    default: lk = l22; x11 = 0; x12 = 0; x22 = l22; break;
  }
}

constexpr int kPostMaxKids = 4;  // sibships up to this size take the pass that looks every kid's likelihoods up once

#ifndef PM_POST_MINB
#define PM_POST_MINB 4   // measured on 200 families x 5 --in_vcf: 1 (254 registers) 15.3, 3 (168) 17.2, 4 (128) 18.1 M records/s
#endif
template <bool CALLS, bool DN>
__global__ void __launch_bounds__(128, DN ? 1 : PM_POST_MINB) k_post(const DevRun *__restrict__ run, const pm_site_hdr *__restrict__ hdr,
                                              const uint4 *__restrict__ recs_all, const pm_site_result *__restrict__ res_all,
                                              const uint32_t *__restrict__ emit_sites, const uint32_t *__restrict__ n_emit_ptr,
                                              size_t res_cap, pm_site_result *__restrict__ res_out,
                                              pm_person_result *__restrict__ person_out, uint16_t *__restrict__ calls_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PostSmem *sm = reinterpret_cast<PostSmem *>(smem_raw);
  load_tables(run, &sm->t);
  __syncthreads();
  const uint32_t n_emit = *n_emit_ptr;
  const size_t n_rows = n_emit < res_cap ? n_emit : res_cap;
  const size_t total = n_rows * (size_t)run->n_fam;
  const int np = run->n_person;
  const double *lut = sm->t.lut, *mut = sm->t.mut;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
    const size_t row = w / run->n_fam;
    const int fi = (int)(w % run->n_fam);
    const uint32_t s = emit_sites[row];
    if ((run->site_filter == 1 && hdr[s].chr_class != PM_CHR_AUTO) || (run->site_filter == 2 && hdr[s].chr_class == PM_CHR_AUTO)) continue;
    pm_site_result r = res_all[s];
    const uint4 *recs = recs_all + (size_t)s * np;
    const PersonSink<CALLS> out{CALLS ? nullptr : person_out + row * (size_t)np, CALLS ? calls_out + row * (size_t)np : nullptr};
    const DevFam f = run->fams[fi];
    if (r.status != PM_SITE_EMITTED) {  // PM_OUT_ALL rows of sites that print nothing
      if (fi == 0) res_out[row] = r;
      for (int j = 0; j < f.size; j++) out.zero(f.first + j);
      continue;
    }
    const int a1 = r.allele1, a2 = r.allele2;
    const int g11 = geno_index(a1, a1), g12 = geno_index(a1, a2), g22 = geno_index(a2, a2);
    const bool mono = (r.flags & PM_FLAG_MONO) != 0;
    const bool dn = DN && run->denovo != 0 && !run->vcf_mode;  // (instances for runs without --denovo leave the ten-state code out)
    // frequency the posteriors are taken at (main:576-587)
    const double freq = mono ? (dn ? 1.0 : 1.0 - run->theta) : r.freq;
    const double q = 1.0 - freq;
    const int cls = hdr[s].chr_class;
    const bool nonauto = cls != PM_CHR_AUTO;

    if (f.kind == 0) {  // CalcPostProb_SinglePerson, NucFam:754-795
      for (int j = 0; j < f.size; j++) {
        uint4 rec = recs[f.first + j];
        double pr0 = freq * freq, pr1 = freq * q * 2, pr2 = q * q;
        const int sex = nonauto ? run->sex[f.first + j] : 0;
        if (cls == PM_CHR_MT || ((cls == PM_CHR_X || cls == PM_CHR_Y) && sex == 1)) { pr0 = freq; pr1 = 0.0; pr2 = q; }
        else if (cls == PM_CHR_Y) { pr0 = pr1 = pr2 = 1.0; }
        double m11 = lut[rec_lk(rec, g11)] * pr0;
        double m12 = lut[rec_lk(rec, g12)] * pr1;
        double m22 = lut[rec_lk(rec, g22)] * pr2;
        double sum = m11 + m12 + m22;
        if (sum == 0 || (cls == PM_CHR_Y && sex == 2)) out.put3(f.first + j, 0, 0, 0, best3(m11, m12, m22));  // NucFam:781, 788
        else out.put3(f.first + j, m11 / sum, m12 / sum, m22 / sum, best3(m11, m12, m22));
      }
    } else if (f.kind == 1) {  // nuclear: NucFam:590-752
      const int nk = f.size - 2;
      double C[9], pp[9], pm9[9];
      const bool na = nonauto && !dn;  // the --denovo nuclear code has no chrX / chrY / MT rules
      // parent-pair prior: HW when nFam>1 (or isMono / freq==1 under --denovo), else the fixed table
      bool hw = run->n_fam > 1 || (dn ? freq == 1.0 : mono);
      if (!na) {
        unit_conditionals(recs, f.first, nk, g11, g12, g22, dn, lut, mut, C);
        if (hw) parent_priors(freq, pp); else single_trio_priors(pp);
      } else {
        // CalcParentMarginal runs before this family's loop assigns `sex` (NucFam:606 vs 610): every kid is given the
        // sex of the LAST member of the previous family, family 0 that of the last person of the previous emitted
        // site's last family -- or the initial 0 in the first CalcPostProb of the process (PM_HDR_FIRST_POSTPROB).
        const int ks = fi > 0 ? run->sex[f.first - 1] : ((hdr[s].reserved & PM_HDR_FIRST_POSTPROB) ? 0 : run->sex[np - 1]);
        unit_conditionals_nonauto(recs, f.first, nk, g11, g12, g22, cls, ks, lut, C);
        if (hw) parent_priors_nonauto(cls, freq, pp); else single_trio_priors(pp);
      }
      for (int j = 0; j < 9; j++) pm9[j] = C[j] * pp[j];
      {
        double p11 = pm9[0] + pm9[1] + pm9[2], p12 = pm9[3] + pm9[4] + pm9[5], p22 = pm9[6] + pm9[7] + pm9[8];
        double sum = p11 + p12 + p22;
        if (sum == 0) out.put3(f.first, 0, 0, 0, best3(p11, p12, p22));
        else out.put3(f.first, p11 / sum, p12 / sum, p22 / sum, best3(p11, p12, p22));
        p11 = pm9[0] + pm9[3] + pm9[6]; p12 = pm9[1] + pm9[4] + pm9[7]; p22 = pm9[2] + pm9[5] + pm9[8];
        sum = p11 + p12 + p22;
        if (sum == 0) out.put3(f.first + 1, 0, 0, 0, best3(p11, p12, p22));
        else out.put3(f.first + 1, p11 / sum, p12 / sum, p22 / sum, best3(p11, p12, p22));
      }
      // parentGLF * parentPrior per configuration (NucFam:815-823)
      double w9[9];
      {
        uint4 rf = recs[f.first], rm = recs[f.first + 1];
        double fl[3] = {lut[rec_lk(rf, g11)], lut[rec_lk(rf, g12)], lut[rec_lk(rf, g22)]};
        double ml[3] = {lut[rec_lk(rm, g11)], lut[rec_lk(rm, g12)], lut[rec_lk(rm, g22)]};
        if (na) {  // NucFam:1049-1051
          fl[1] = 0.0;
          if (cls == PM_CHR_Y) ml[0] = ml[1] = ml[2] = 1.0;
          if (cls == PM_CHR_MT) ml[1] = 0.0;
        }
        for (int j = 0; j < 9; j++) w9[j] = (fl[j / 3] * ml[j % 3]) * pp[j];
      }
      if (!dn && nk <= kPostMaxKids) {
        // KidJointGenoLikelihood for every kid of a small sibship in one pass: the reference (and the loop below) looks the
        // nk kids' likelihoods up and runs likelihoodKidGenotype nk x 9 x nk times; here every kid's three likelihoods are
        // fetched once and its (lk, x11, x12, x22) taken once per configuration.  Same products and sums in the same order.
        double kl[kPostMaxKids][3];
        int ksex[kPostMaxKids];
#pragma unroll
        for (int kk = 0; kk < kPostMaxKids; kk++) {
          kl[kk][0] = kl[kk][1] = kl[kk][2] = 1.0; ksex[kk] = 0;
          if (kk < nk) {
            const uint4 rk = recs[f.first + 2 + kk];
            kl[kk][0] = lut[rec_lk(rk, g11)]; kl[kk][1] = lut[rec_lk(rk, g12)]; kl[kk][2] = lut[rec_lk(rk, g22)];
            if (na) ksex[kk] = run->sex[f.first + 2 + kk];
          }
        }
        double J[kPostMaxKids][3];
#pragma unroll
        for (int cfg = 0; cfg < 9; cfg++) {
          double lkv[kPostMaxKids], xv[kPostMaxKids][3];
#pragma unroll
          for (int kk = 0; kk < kPostMaxKids; kk++)
            kid_cfg(cfg, na, cls, ksex[kk], kl[kk][0], kl[kk][1], kl[kk][2], lkv[kk], xv[kk][0], xv[kk][1], xv[kk][2]);
#pragma unroll
          for (int kid = 0; kid < kPostMaxKids; kid++) {
            double G[3] = {1.0, 1.0, 1.0};
#pragma unroll
            for (int kk = 0; kk < kPostMaxKids; kk++) {
              if (kk >= nk) continue;
              if (kk != kid) { G[0] *= lkv[kk]; G[1] *= lkv[kk]; G[2] *= lkv[kk]; }
              else { G[0] *= xv[kk][0]; G[1] *= xv[kk][1]; G[2] *= xv[kk][2]; }
            }
#pragma unroll
            for (int t = 0; t < 3; t++) J[kid][t] = cfg == 0 ? G[t] * w9[cfg] : J[kid][t] + G[t] * w9[cfg];
          }
        }
#pragma unroll
        for (int kid = 0; kid < kPostMaxKids; kid++) {
          if (kid >= nk) continue;
          const double sum = J[kid][0] + J[kid][1] + J[kid][2];
          double p0 = 0, p1 = 0, p2 = 0;
          if (sum != 0.0) { p0 = J[kid][0] / sum; p1 = J[kid][1] / sum; p2 = J[kid][2] / sum; }
          out.put3(f.first + 2 + kid, p0, p1, p2, best3(p0, p1, p2));
        }
      } else
      for (int kid = 0; kid < nk; kid++) {
        const int o = f.first + 2 + kid;
        if (!dn) {
          // KidJointGenoLikelihood + likelihoodKidGenotype, NucFam:798-835, 1334-1443
          double J[3] = {0, 0, 0};
          for (int cfg = 0; cfg < 9; cfg++) {
            double G[3] = {1.0, 1.0, 1.0};
            for (int kk = 0; kk < nk; kk++) {
              uint4 rk = recs[f.first + 2 + kk];
              double l11 = lut[rec_lk(rk, g11)], l12 = lut[rec_lk(rk, g12)], l22 = lut[rec_lk(rk, g22)];
              double lk, x11, x12, x22;
              kid_cfg(cfg, na, cls, na ? run->sex[f.first + 2 + kk] : 0, l11, l12, l22, lk, x11, x12, x22);
              if (kk != kid) { G[0] *= lk; G[1] *= lk; G[2] *= lk; }
              else { G[0] *= x11; G[1] *= x12; G[2] *= x22; }
            }
            for (int t = 0; t < 3; t++) J[t] = cfg == 0 ? G[t] * w9[cfg] : J[t] + G[t] * w9[cfg];
          }
          double sum = J[0] + J[1] + J[2];
          double p0 = 0, p1 = 0, p2 = 0;
          if (sum != 0.0) { p0 = J[0] / sum; p1 = J[1] / sum; p2 = J[2] / sum; }
          out.put3(o, p0, p1, p2, best3(p0, p1, p2));
        } else {
          // KidJointGenoLikelihood_denovo, NucFam:838-868, 1446-1551
          double geno[10];
          for (int g = 0; g < 10; g++) geno[g] = 0.0;
          const double *M1 = mut + g11 * 10, *M2 = mut + g12 * 10, *M3 = mut + g22 * 10;
          for (int cfg = 0; cfg < 9; cfg++) {
            double lkg[10];
            for (int g = 0; g < 10; g++) lkg[g] = 1.0;
            for (int kk = 0; kk < nk; kk++) {
              uint4 rk = recs[f.first + 2 + kk];
              if (kk != kid) {
                double d11 = 0, d12 = 0, d22 = 0;
                for (int g = 0; g < 10; g++) {
                  double l = lut[rec_lk(rk, g)];
                  d11 += M1[g] * l; d12 += M2[g] * l; d22 += M3[g] * l;
                }
                double lk;
                switch (cfg) {
                  case 0: lk = d11; break;
                  case 1: case 3: lk = 0.5 * (d11 + d12); break;
                  case 2: case 6: lk = d12; break;
                  case 4: lk = 0.25 * d11 + 0.5 * d12 + 0.25 * d22; break;
                  case 5: case 7: lk = 0.5 * (d12 + d22); break;
                  default: lk = d22; break;
                }
                for (int g = 0; g < 10; g++) lkg[g] *= lk;
              } else {
                for (int g = 0; g < 10; g++) {
                  double l = lut[rec_lk(rk, g)], wgt;
                  switch (cfg) {
                    case 0: wgt = M1[g]; break;
                    case 1: case 3: wgt = 0.5 * M1[g] + 0.5 * M2[g]; break;
                    case 2: case 6: wgt = M2[g]; break;
                    case 4: wgt = 0.25 * M1[g] + 0.5 * M2[g] + 0.25 * M3[g]; break;
                    case 5: case 7: wgt = 0.5 * M2[g] + 0.5 * M3[g]; break;
                    default: wgt = M3[g]; break;
                  }
                  lkg[g] *= wgt * l;
                }
              }
            }
            for (int g = 0; g < 10; g++) geno[g] += lkg[g] * w9[cfg];
          }
          double sum = 0.0;
          for (int g = 0; g < 10; g++) sum += geno[g];
          double mx = 0.0;
          int best = 0;
          for (int g = 0; g < 10; g++) {
            double pg = sum == 0.0 ? 0.0 : geno[g] / sum;
            geno[g] = pg;
            if (mx < pg) { mx = pg; best = g; }
          }
          out.put10(o, geno, best);
        }
      }
    }  // (extended families: k_post_es, a warp per person)
    if (fi == 0) {
      // CalculateAB (NucFam:1006-1039), only printed by the non-de-novo writer on autosomes
      if (nonauto) {
        r.ab = 0.0;  // not computed and not printed there (NucFam:1791-1800)
      } else if (!dn && !run->vcf_mode) {
        r.ab = 0.5;  // k_post_ab, launched right behind this kernel, puts the allele balance here (a sum over ALL persons of the site)
      } else {
        r.ab = 0.5;
      }
      res_out[row] = r;
    }
  }
}


// Genotype posteriors of extended-family members (FLSeq:140-216): every person's genotypes are pinned in turn and the
// family re-peeled -- 3 peels per person bi-allelic, 10 under --denovo: 200 ten-state peels for a 20-member pedigree.
// On one thread per (row, family) (round 1) that chain took ~4 ms however few rows a batch emitted (a third of a CEPH
// --denovo step); here a warp owns one (row, family, person) and lane g < A peels with genotype g pinned.  The same
// es_likelihood<A> per genotype, the values summed in genotype order by lane 0: the same bits as the serial loop.
template <int A, bool CALLS>
__global__ void __launch_bounds__(128) k_post_es(const DevRun *__restrict__ run, const pm_site_hdr *__restrict__ hdr,
                                                 const uint4 *__restrict__ recs_all, const pm_site_result *__restrict__ res_all,
                                                 const uint32_t *__restrict__ emit_sites, const uint32_t *__restrict__ n_emit_ptr,
                                                 size_t res_cap, pm_person_result *__restrict__ person_out, uint16_t *__restrict__ calls_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PostSmem *sm = reinterpret_cast<PostSmem *>(smem_raw);
  load_tables(run, &sm->t);
  __syncthreads();
  constexpr bool dn = A == 10;
  if ((run->denovo != 0 && run->vcf_mode == 0) != dn || run->n_es == 0) return;
  const uint32_t n_emit = *n_emit_ptr;
  const size_t n_rows = n_emit < res_cap ? n_emit : res_cap;
  const int np = run->n_person, lane = threadIdx.x & 31;
  const size_t per_row = (size_t)run->n_es * kMaxEsPersons;
  const size_t total = n_rows * per_row;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const double *lut = sm->t.lut, *mut = sm->t.mut;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += n_warps) {
    const size_t row = w / per_row;
    const int e = (int)((w % per_row) / kMaxEsPersons), j = (int)(w % kMaxEsPersons);
    const DevFam f = run->fams[run->es_fams[e]];
    if (j >= f.size) continue;
    const uint32_t s = emit_sites[row];
    const int cls = hdr[s].chr_class;
    if ((run->site_filter == 1 && cls != PM_CHR_AUTO) || (run->site_filter == 2 && cls == PM_CHR_AUTO)) continue;
    const pm_site_result *r = res_all + s;
    if (r->status != PM_SITE_EMITTED) continue;  // (k_post has zeroed the row's persons)
    const PersonSink<CALLS> out{CALLS ? nullptr : person_out + row * (size_t)np, CALLS ? calls_out + row * (size_t)np : nullptr};
    const int o = f.first + j;
    if (!dn && cls == PM_CHR_Y && run->sex[o] == 2) {  // FLSeq:181-188
      if (lane == 0) out.put3(o, 0, 0, 0, 0);
      continue;
    }
    const int a1 = r->allele1, a2 = r->allele2;
    const int g11 = geno_index(a1, a1), g12 = geno_index(a1, a2), g22 = geno_index(a2, a2);
    // frequency the posteriors are taken at (main:576-587)
    const double freq = (r->flags & PM_FLAG_MONO) ? (dn ? 1.0 : 1.0 - run->theta) : r->freq;
    const uint4 *recs = recs_all + (size_t)s * np;
    const int pinned = dn ? lane : (lane == 0 ? g11 : (lane == 1 ? g12 : g22));
    double mine = 0.0;
    if (lane < A) mine = es_likelihood<A>(run, f, recs, g11, g12, g22, dn, freq, lut, mut, j, pinned, cls);
    double lk[A];
#pragma unroll
    for (int g = 0; g < A; g++) lk[g] = __shfl_sync(0xffffffffu, mine, g);
    if (lane != 0) continue;
    if constexpr (!dn) {
      const double sum = lk[0] + lk[1] + lk[2];
      if (sum == 0) out.put3(o, 0, 0, 0, best3(lk[0], lk[1], lk[2]));
      else out.put3(o, lk[0] / sum, lk[1] / sum, lk[2] / sum, best3(lk[0], lk[1], lk[2]));
    } else {
      double sum = 0.0;
      for (int g = 0; g < A; g++) sum += lk[g];
      double mx = 0.0;
      int best = 0;
      for (int g = 0; g < A; g++) {
        if (mx < lk[g]) { mx = lk[g]; best = g; }
        lk[g] = sum == 0 ? 0.0 : lk[g] / sum;
      }
      out.put10(o, lk, best);
    }
  }
}

// CalculateAB (NucFam:1006-1039) for the emitted autosomal rows of a run without --denovo: one block per row, the persons
// spread over its threads (the loop is 3,000 divisions long for 1,000 trios: on one thread of k_post it took longer than
// everything else in the batch).  Per-person terms exactly as the reference's; their sum is taken as a tree (the printed
// digits do not see the difference in the last bits).
__global__ void __launch_bounds__(128) k_post_ab(const DevRun *__restrict__ run, const pm_site_hdr *__restrict__ hdr,
                                                 const uint4 *__restrict__ recs_all, const pm_site_result *__restrict__ res_all,
                                                 const uint32_t *__restrict__ emit_sites, const uint32_t *__restrict__ n_emit_ptr,
                                                 size_t res_cap, pm_site_result *__restrict__ res_out) {
  __shared__ double s_lut[256];
  __shared__ double s_red[2][4];
  if (run->denovo != 0 || run->vcf_mode != 0) return;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = run->lut[i];
  __syncthreads();
  const uint32_t n_emit = *n_emit_ptr;
  const size_t n_rows = n_emit < res_cap ? n_emit : res_cap;
  const int np = run->n_person;
  for (size_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const uint32_t s = emit_sites[row];
    const int cls = hdr[s].chr_class;
    if ((run->site_filter == 1 && cls != PM_CHR_AUTO) || (run->site_filter == 2 && cls == PM_CHR_AUTO)) continue;
    const pm_site_result *r = res_all + s;
    if (r->status != PM_SITE_EMITTED || cls != PM_CHR_AUTO) continue;
    const int a1 = r->allele1, a2 = r->allele2;
    const int g11 = geno_index(a1, a1), g12 = geno_index(a1, a2), g22 = geno_index(a2, a2);
    const double f0 = r->freq;
    const double p11 = f0 * f0, p12 = 2 * f0 * (1 - f0), p22 = (1 - f0) * (1 - f0);
    const uint4 *recs = recs_all + (size_t)s * np;
    double A = 0.0, Bsum = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
      const uint4 rec = recs[i];
      const int depth = rec_depth(rec);
      const int u11 = rec_lk(rec, g11), u12 = rec_lk(rec, g12), u22 = rec_lk(rec, g22);
      const double l11 = s_lut[u11], l12 = s_lut[u12], l22 = s_lut[u22];
      const double phet = (p12 * l12) / (p11 * l11 + p12 * l12 + p22 * l22);
      if (phet > 1e-05 && depth > 0) {
        int scale = u22 + u11 - 2 * u12 + 6 * depth;
        const int minimum = abs(u22 - u11);
        if (scale < 4) scale = 4;
        if (scale < minimum) scale = minimum;
        const int nref = (int)(0.5 * depth * (1 + (u22 - u11) / (scale + 1e-30)));
        A += phet * nref;
        Bsum += phet * depth;
      }
    }
    for (int o = 16; o > 0; o >>= 1) { A += __shfl_xor_sync(0xffffffffu, A, o); Bsum += __shfl_xor_sync(0xffffffffu, Bsum, o); }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = A; s_red[1][threadIdx.x >> 5] = Bsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += s_red[0][w]; b += s_red[1][w]; }
      res_out[row].ab = (0.05 + a) / (0.1 + b);
    }
    __syncthreads();
  }
}

cudaError_t launch_post(const DevRun *d_run, int n_fam, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                        const pm_site_result *d_res_all, const uint32_t *d_emit_sites, const uint32_t *d_n_emit,
                        size_t max_rows, size_t res_cap, pm_site_result *d_res_out, pm_person_result *d_person_out,
                        uint16_t *d_calls_out, bool has_es, bool ten_state, int sm_count, bool with_ab, cudaStream_t stream) {
  if (max_rows == 0) return cudaSuccess;
  static_assert(sizeof(PostSmem) <= 48 * 1024, "k_post's tables fit in the default dynamic shared memory limit");
  size_t want = (max_rows * (size_t)n_fam + 127) / 128;
  size_t cap = (size_t)sm_count * 16;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  if (grid == 0) grid = 1;
#define PM_POST(CALLS_, DN_) \
  k_post<CALLS_, DN_><<<grid, 128, sizeof(PostSmem), stream>>>(d_run, d_hdr, d_recs, d_res_all, d_emit_sites, d_n_emit, res_cap, d_res_out, d_person_out, d_calls_out)
  if (d_calls_out) { if (ten_state) PM_POST(true, true); else PM_POST(true, false); }
  else             { if (ten_state) PM_POST(false, true); else PM_POST(false, false); }
#undef PM_POST
  if (has_es) {  // the extended families' members: a warp per (row, family, person)
    size_t warps = max_rows * 64;  // enough to start with; the kernel strides over the rest
    const size_t es_cap = (size_t)sm_count * 8 * 4;
    if (warps > es_cap) warps = es_cap;
    const unsigned es_grid = (unsigned)((warps + 3) / 4);
#define PM_POST_ES(A_, CALLS_) \
  k_post_es<A_, CALLS_><<<es_grid, 128, sizeof(PostSmem), stream>>>(d_run, d_hdr, d_recs, d_res_all, d_emit_sites, d_n_emit, res_cap, d_person_out, d_calls_out)
    if (ten_state) { if (d_calls_out) PM_POST_ES(10, true); else PM_POST_ES(10, false); }
    else           { if (d_calls_out) PM_POST_ES(3, true); else PM_POST_ES(3, false); }
#undef PM_POST_ES
  }
  if (!with_ab) return cudaGetLastError();
  const size_t ab_cap = (size_t)sm_count * 8;
  k_post_ab<<<(unsigned)(max_rows < ab_cap ? max_rows : ab_cap), 128, 0, stream>>>(d_run, d_hdr, d_recs, d_res_all, d_emit_sites, d_n_emit, res_cap, d_res_out);
  return cudaGetLastError();
}

}  // namespace pm
