// pm_device.cuh — device-side math shared by the site kernels (sm_100a).
//
// What is computed follows the reference's per-site likelihood engine (file:line cited per function;
// "NucFam" = src/NucFamGenotypeLikelihood.cpp, "FLSeq" = src/FamilyLikelihoodSeq.cpp,
// "ES" = src/FamilyLikelihoodES.cpp, "Gold" = core/MathGold.cpp, "main" = src/main.cpp).  How it is
// computed is our own:
//   * a nuclear family's likelihood  L_f(p) = sum_j prior_j(p) * C_fj  (NucFam:941-1132) is split into
//     the nine p-independent coefficients C_fj (built once per site and hypothesis) and a quartic form
//     L_f(p) = sum_a B_fa p^a q^(4-a) with five coefficients, because prior_j(p) = c_j p^a q^(4-a);
//     unrelated founders are the same form with (p+q)^2 folded in.  One objective evaluation is then
//     5 FMAs per family instead of ~100 flops, and the sum of log10 over families becomes one log10 of
//     a running product with the exponent tracked in integers.
//   * Brent (Gold:81-177) is a resumable state machine so that one thread can drive it while a whole
//     block evaluates the objective.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "polymutt_b200.h"

namespace pm {

constexpr int kMaxEsPersons = 64;  // largest extended family the per-thread peel workspace holds (local memory; only the used part is touched)
constexpr int kMaxMp = 16;         // marriage-partial slots per extended family
constexpr double kLog10_2 = 0.30102999566398119521;

// One resolved peel step (ES:990-1057): which marriage-partial slot it touches is static, so the
// std::map lookups of the reference are done once on the host.
struct DevStep {
  int8_t type;   // PM_PEEL_*
  int8_t from0, from1, to0, to1;
  int8_t mp;     // marriage partial slot, -1 = none found (exact (first,second) key match, as the reference)
  int8_t flag;   // type 1: 1 = slot is created (all ones) by this step; type 2: 1 = from is the father (fa2mo)
  int8_t pad;
};

struct DevFam {       // one pedigree family, VCF column order
  int32_t first;      // column of member 0
  int16_t size, founders;
  int8_t kind;        // 0 = founders only, 1 = nuclear, 2 = extended (Elston-Stewart)
  int8_t n_mp;
  int16_t n_steps;
  int32_t step_first; // index into the run-wide concatenation of all families' peel steps (can pass 32,767)
};
static_assert(sizeof(DevFam) == 16, "DevFam is copied to the device as 16-byte records");

struct DevUnit {      // one factor of the objective that has the quartic form
  int32_t first;      // column of the first member
  int32_t nkids;      // -1 = a single unrelated founder; >= 0 = nuclear family with that many kids
  int32_t kid0;       // number of kids in the units before this one (slot of its first kid in the per-site kid table)
  int32_t sex;        // single founder: that person's sex (1 male, 2 female), read on chrX / chrY only
};

// Everything the kernels need about the run; lives in global memory, hot tables are copied to smem.
struct DevRun {
  double lut[256];            // 10^(-i/10)   (core/BaseQualityHelper.cpp:12-13), host-computed
  double mut[100];            // genotype mutation matrix (src/MutationModel.cpp:46-90), host-computed
  double log_inv[128];        // 1/c_i, c_i = 1 + (i+0.5)/128: table-driven log10 of a mantissa in [1,2)
  double log_tab[128];        // -log10(log_inv[i]) (computed in long double on the host)
  // host-computed log10 constants (same libm as the reference)
  double log_1m_prior, log_prior_ts, log_prior_tv, log_prior_other, log_prior_23, log_prior_16, log_min_llr;
  // the same six per chromosome class PM_CHR_* (SetPolyPrior_chrX/_chrY/_MT, NucFam:256-293):
  // [cls][0..5] = log_1m_prior, log_prior_ts, log_prior_tv, log_prior_other, log_prior_23, log_prior_16
  double cls_log[4][6];
  double vcf_log_ts, vcf_log_tv, vcf_log_indel;  // VCF mode: log10(2/3), log10(1/6), log10(prior)  (src/PedVCF.cpp:143-150)
  double theta, posterior_cutoff, precision, denovo_min_llr, min_ps;
  int32_t min_map_quality, min_total_depth, max_total_depth;
  int32_t denovo, force_call, out_all_sites;
  int32_t site_filter;        // 0 = every site; 1 = autosomal sites only, 2 = chrX / chrY / MT only: VCF input keeps two
                              // descriptions of the pedigree and each one leaves the other's records untouched
  int32_t vcf_mode;           // 1 = records come from a VCF (src/PedVCF.cpp:116-163): one hypothesis (REF, ALT), QUAL formula
  int32_t n_person, n_fam, n_units, n_es, n_kids;
  int32_t use_brent;          // nFam>1 || !nuclear  (FLSeq:94)
  unsigned long long *counters;  // [4] hypotheses, evaluations, sites evaluated, sites emitted
  const DevFam *fams;
  const DevUnit *units;
  const int32_t *es_fams;     // indices into fams[] of the extended families
  const DevStep *steps;
  const uint8_t *sex;         // per column: 1 male, 2 female (chrX / chrY rules)
};

// ---- small helpers ----------------------------------------------------------------------------
__host__ __device__ inline int geno_index(int b1, int b2) {  // core/glfHandler.h:102-106
  return b1 < b2 ? (b1 - 1) * (10 - b1) / 2 + (b2 - b1) : (b2 - 1) * (10 - b2) / 2 + (b1 - b2);
}
__host__ __device__ inline int poly_ts(int r) { return ((r - 1) ^ 2) + 1; }    // src/PedigreeGLF.h:15-27
__host__ __device__ inline int poly_tvs1(int r) { return (r & 1) ? 2 : 1; }    // :28-40
__host__ __device__ inline int poly_tvs2(int r) { return (r & 1) ? 4 : 3; }    // :41-53

__device__ __forceinline__ uint32_t rec_lk(const uint4 &r, int g) {
  uint32_t w = g < 4 ? r.x : (g < 8 ? r.y : r.z);
  return (w >> ((g & 3) * 8)) & 0xffu;
}
__device__ __forceinline__ int rec_depth(const uint4 &r) { return (int)((r.z >> 16) | ((r.w & 0xffu) << 16)); }
__device__ __forceinline__ int rec_mapq(const uint4 &r) { return (int)((r.w >> 8) & 0xffu); }

// ---- quartic coefficients of one unit -----------------------------------------------------------
// C_j = likelihoodKids(j) * lF * lM  (NucFam:1041-1132, 1184-1312);  returns B with
// L(p) = B[4] p^4 + B[3] p^3 q + B[2] p^2 q^2 + B[1] p q^3 + B[0] q^4.
// `C` (optional) receives the nine parentConditional values (bit-identical to the reference's).
template <typename RecPtr>
__device__ __forceinline__ void unit_conditionals(RecPtr recs, int first, int nkids, int g11, int g12, int g22,
                                                  bool denovo, const double *__restrict__ lut,
                                                  const double *__restrict__ mut, double C[9]) {
  uint4 rf = recs[first], rm = recs[first + 1];
  double f11 = lut[rec_lk(rf, g11)], f12 = lut[rec_lk(rf, g12)], f22 = lut[rec_lk(rf, g22)];
  double m11 = lut[rec_lk(rm, g11)], m12 = lut[rec_lk(rm, g12)], m22 = lut[rec_lk(rm, g22)];
  double p0 = 1.0, p1 = 1.0, p2 = 1.0, p4 = 1.0, p5 = 1.0, p8 = 1.0;
  for (int k = 0; k < nkids; k++) {
    uint4 rk = recs[first + 2 + k];
    double d11, d12, d22;
    if (denovo) {  // CalcDenovoMutLk, NucFam:1553-1562
      // ten independent look-ups first, then the three dot products as two half-sums each: short dependency
      // chains matter more than instruction count here (few resident warps per SM)
      double l[10];
#pragma unroll
      for (int g = 0; g < 10; g++) l[g] = lut[rec_lk(rk, g)];
      const double *r11 = mut + g11 * 10, *r12 = mut + g12 * 10, *r22 = mut + g22 * 10;
      double a11 = 0.0, b11 = 0.0, a12 = 0.0, b12 = 0.0, a22 = 0.0, b22 = 0.0;
#pragma unroll
      for (int g = 0; g < 5; g++) {
        a11 += r11[g] * l[g]; b11 += r11[g + 5] * l[g + 5];
        a12 += r12[g] * l[g]; b12 += r12[g + 5] * l[g + 5];
        a22 += r22[g] * l[g]; b22 += r22[g + 5] * l[g + 5];
      }
      d11 = a11 + b11; d12 = a12 + b12; d22 = a22 + b22;
    } else {
      d11 = lut[rec_lk(rk, g11)]; d12 = lut[rec_lk(rk, g12)]; d22 = lut[rec_lk(rk, g22)];
    }
    // likelihoodONEKid{,_denovo}, NucFam:1202-1296 (autosome)
    p0 *= d11;
    p1 *= 0.5 * (d11 + d12);
    p2 *= d12;
    p4 *= 0.25 * d11 + 0.5 * d12 + 0.25 * d22;
    p5 *= 0.5 * (d12 + d22);
    p8 *= d22;
  }
  C[0] = p0 * (f11 * m11); C[1] = p1 * (f11 * m12); C[2] = p2 * (f11 * m22);
  C[3] = p1 * (f12 * m11); C[4] = p4 * (f12 * m12); C[5] = p5 * (f12 * m22);
  C[6] = p2 * (f22 * m11); C[7] = p5 * (f22 * m12); C[8] = p8 * (f22 * m22);
}

__device__ __forceinline__ void quartic_from_conditionals(const double C[9], double B[5]) {
  B[4] = C[0];
  B[3] = 2.0 * (C[1] + C[3]);
  B[2] = C[2] + 4.0 * C[4] + C[6];
  B[1] = 2.0 * (C[5] + C[7]);
  B[0] = C[8];
}

// ---- chrX / chrY / MT (bi-allelic model only; the --denovo nuclear code has no such rules) -------
// likelihoodONEKid on a non-autosome (NucFam:1202-1264) for the six parent configurations whose prior is not
// zero there (father "homozygous"): q = (cfg0, cfg1, cfg2, cfg6, cfg7, cfg8).  `ks` is NOT the kid's sex but the
// object's stale member `sex`: 0 for the hypothesis objects, see k_post for famlk[0].
__device__ __forceinline__ void onekid_nonauto(int cls, int ks, double l11, double l12, double l22, double q[6]) {
  const bool male = ks == 1, female = ks == 2;
  if (cls == PM_CHR_X) {
    q[0] = l11; q[5] = l22;
    q[1] = male ? 0.5 * (l11 + l22) : 0.5 * (l11 + l12);
    q[2] = male ? l22 : l12;
    q[3] = male ? l11 : l12;
    q[4] = male ? 0.5 * (l11 + l22) : 0.5 * (l12 + l22);
  } else if (cls == PM_CHR_Y) {
    q[0] = female ? 1.0 : l11; q[5] = female ? 1.0 : l22;
    q[1] = male ? l11 : 1.0; q[2] = male ? l11 : 1.0;
    q[3] = male ? l22 : 1.0; q[4] = male ? l22 : 1.0;
  } else {
    q[0] = l11; q[5] = l22;
    q[1] = 0.5 * (l11 + l22); q[2] = l22; q[3] = l11; q[4] = 0.5 * (l11 + l22);
  }
}
// parentConditional on a non-autosome: CalcParentMarginal's parentGLF edits (NucFam:1049-1051) and the kid
// products above; C[3..5] (father heterozygous) are zero.
template <typename RecPtr>
__device__ __forceinline__ void unit_conditionals_nonauto(RecPtr recs, int first, int nkids, int g11, int g12, int g22, int cls, int ks,
                                                          const double *__restrict__ lut, double C[9]) {
  uint4 rf = recs[first], rm = recs[first + 1];
  double f11 = lut[rec_lk(rf, g11)], f22 = lut[rec_lk(rf, g22)];
  double m11 = lut[rec_lk(rm, g11)], m12 = lut[rec_lk(rm, g12)], m22 = lut[rec_lk(rm, g22)];
  if (cls == PM_CHR_Y) m11 = m12 = m22 = 1.0;
  if (cls == PM_CHR_MT) m12 = 0.0;
  double p[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
  for (int k = 0; k < nkids; k++) {
    uint4 rk = recs[first + 2 + k];
    double q[6];
    onekid_nonauto(cls, ks, lut[rec_lk(rk, g11)], lut[rec_lk(rk, g12)], lut[rec_lk(rk, g22)], q);
#pragma unroll
    for (int j = 0; j < 6; j++) p[j] *= q[j];
  }
  C[0] = p[0] * (f11 * m11); C[1] = p[1] * (f11 * m12); C[2] = p[2] * (f11 * m22);
  C[3] = 0.0; C[4] = 0.0; C[5] = 0.0;
  C[6] = p[3] * (f22 * m11); C[7] = p[4] * (f22 * m12); C[8] = p[5] * (f22 * m22);
}
// SetParentPrior on a non-autosome (NucFam:333-366)
__device__ __forceinline__ void parent_priors_nonauto(int cls, double freq, double pp[9]) {
  const double q = 1.0 - freq;
  pp[3] = pp[4] = pp[5] = 0.0;
  if (cls == PM_CHR_X) {
    pp[0] = freq * freq * freq; pp[1] = freq * freq * q * 2; pp[2] = freq * q * q;
    pp[6] = q * freq * freq; pp[7] = q * freq * q * 2; pp[8] = q * q * q;
  } else if (cls == PM_CHR_Y) {
    pp[0] = pp[1] = pp[2] = freq; pp[6] = pp[7] = pp[8] = q;
  } else {
    pp[0] = freq * freq; pp[1] = 0.0; pp[2] = freq * q; pp[6] = q * freq; pp[7] = 0.0; pp[8] = q * q;
  }
}
// L(p) = sum_j C_j prior_j(p) has degree 3 (X), 1 (Y) or 2 (MT) there; multiplied by the right power of
// (p + q) = 1 it takes the same quartic form as on the autosomes, with non-negative terms only.
__device__ __forceinline__ void quartic_from_conditionals_nonauto(int cls, const double C[9], double B[5]) {
  if (cls == PM_CHR_X) {
    const double c3 = C[0], c2 = 2.0 * C[1] + C[6], c1 = C[2] + 2.0 * C[7], c0 = C[8];
    B[4] = c3; B[3] = c3 + c2; B[2] = c2 + c1; B[1] = c1 + c0; B[0] = c0;
  } else if (cls == PM_CHR_Y) {
    const double d1 = (C[0] + C[1]) + C[2], d0 = (C[6] + C[7]) + C[8];
    B[4] = d1; B[3] = 3.0 * d1 + d0; B[2] = 3.0 * (d1 + d0); B[1] = d1 + 3.0 * d0; B[0] = d0;
  } else {
    const double e2 = C[0], e1 = C[2] + C[6], e0 = C[8];
    B[4] = e2; B[3] = 2.0 * e2 + e1; B[2] = e2 + 2.0 * e1 + e0; B[1] = e1 + 2.0 * e0; B[0] = e0;
  }
}
// lkSinglePerson (NucFam:987-1004) as a quartic: haploid l11 p + l22 q times (p+q)^3, a female on Y is the constant 1
__device__ __forceinline__ bool single_person_nonauto(int cls, int sex, double l11, double l22, double B[5]) {
  const bool male = sex == 1;
  if (cls == PM_CHR_X && !male) return false;  // diploid, as on the autosomes
  if (cls == PM_CHR_Y && !male) { B[4] = 1.0; B[3] = 4.0; B[2] = 6.0; B[1] = 4.0; B[0] = 1.0; return true; }
  B[4] = l11; B[3] = 3.0 * l11 + l22; B[2] = 3.0 * (l11 + l22); B[1] = l11 + 3.0 * l22; B[0] = l22;
  return true;
}
// One unit on chromosome class cls != PM_CHR_AUTO.  Under --denovo only the single founders differ from the
// autosomal code (lkSingleFam_denovo -> lkSinglePerson); nuclear families keep CalcParentMarginal_denovo.
template <typename RecPtr>
__device__ __forceinline__ void unit_quartic_nonauto(RecPtr recs, const DevUnit u, int g11, int g12, int g22, bool denovo, int cls, int ks,
                                                     const double *__restrict__ lut, const double *__restrict__ mut, double B[5]);

template <typename RecPtr>
__device__ __forceinline__ void unit_quartic(RecPtr recs, const DevUnit u, int g11, int g12, int g22, bool denovo,
                                             const double *__restrict__ lut, const double *__restrict__ mut,
                                             double B[5]) {
  if (u.nkids < 0) {  // lkSinglePerson, NucFam:987-1004, times (p+q)^2
    uint4 r = recs[u.first];
    double l11 = lut[rec_lk(r, g11)], l12 = lut[rec_lk(r, g12)], l22 = lut[rec_lk(r, g22)];
    B[4] = l11;
    B[3] = 2.0 * (l11 + l12);
    B[2] = l11 + 4.0 * l12 + l22;
    B[1] = 2.0 * (l12 + l22);
    B[0] = l22;
  } else {
    double C[9];
    unit_conditionals(recs, u.first, u.nkids, g11, g12, g22, denovo, lut, mut, C);
    quartic_from_conditionals(C, B);
  }
}

template <typename RecPtr>
__device__ __forceinline__ void unit_quartic_nonauto(RecPtr recs, const DevUnit u, int g11, int g12, int g22, bool denovo, int cls, int ks,
                                                     const double *__restrict__ lut, const double *__restrict__ mut, double B[5]) {
  if (u.nkids < 0) {
    uint4 r = recs[u.first];
    if (!single_person_nonauto(cls, u.sex, lut[rec_lk(r, g11)], lut[rec_lk(r, g22)], B)) unit_quartic(recs, u, g11, g12, g22, denovo, lut, mut, B);
  } else if (denovo) {
    unit_quartic(recs, u, g11, g12, g22, true, lut, mut, B);
  } else {
    double C[9];
    unit_conditionals_nonauto(recs, u.first, u.nkids, g11, g12, g22, cls, ks, lut, C);
    quartic_from_conditionals_nonauto(cls, C, B);
  }
}

struct Monomials { double m4, m3, m2, m1, m0; };
__device__ __forceinline__ Monomials monomials(double p) {
  double q = 1.0 - p, p2 = p * p, q2 = q * q, pq = p * q;
  Monomials m;
  m.m4 = p2 * p2; m.m3 = p2 * pq; m.m2 = p2 * q2; m.m1 = pq * q2; m.m0 = q2 * q2;
  return m;
}
__device__ __forceinline__ double quartic_eval(const double B[5], const Monomials &m) {
  // two independent pairs then one add and one fma: dependency depth 4 instead of the 5 of a straight fma chain
  // (the evaluation is bound by FP64 latency, not throughput)
  const double t1 = fma(B[3], m.m3, B[4] * m.m4), t2 = fma(B[1], m.m1, B[2] * m.m2);
  return fma(B[0], m.m0, t1 + t2);
}

// HW parent-pair priors (NucFam:323-331) and the fixed single-trio table (NucFam:383-394).
__device__ __forceinline__ void parent_priors(double freq, double pp[9]) {
  double q = 1.0 - freq;
  pp[0] = freq * freq * freq * freq;
  pp[1] = freq * freq * freq * q * 2;
  pp[2] = freq * freq * q * q;
  pp[3] = freq * q * 2 * freq * freq;
  pp[4] = freq * q * 2 * freq * q * 2;
  pp[5] = freq * q * 2 * q * q;
  pp[6] = q * q * freq * freq;
  pp[7] = q * q * freq * q * 2;
  pp[8] = q * q * q * q;
}
__device__ __forceinline__ void single_trio_priors(double pp[9]) {
  pp[0] = 0.0; pp[1] = 0.24; pp[2] = 0.04; pp[3] = 0.24; pp[4] = 0.16; pp[5] = 0.08; pp[6] = 0.04; pp[7] = 0.08; pp[8] = 0.12;
}

// ---- running product with integer exponent ------------------------------------------------------
// sum_u log10(L_u) = log10(prod mant) + (sum exp) * log10(2).
struct ProdAcc { double m; int e; };
__device__ __forceinline__ void prod_init(ProdAcc &a) { a.m = 1.0; a.e = 0; }
__device__ __forceinline__ void prod_mul(ProdAcc &a, double x) {
  int hi = __double2hiint(x);
  int ex = (hi >> 20) & 0x7ff;
  if (ex == 0 || ex == 0x7ff || hi < 0) {
    // zero / subnormal / inf / nan / negative: the reference would see log10 of it (-inf, nan).  We keep
    // the objective finite: zero counts as 2^-(2^24); subnormals are rescaled.
    if (!(x > 0.0)) { a.e -= (1 << 24); return; }
    if (ex == 0x7ff) { a.e += (1 << 24); return; }
    x *= 1.3407807929942597e154;  // 2^512
    a.e -= 512;
    hi = __double2hiint(x);
    ex = (hi >> 20) & 0x7ff;
  }
  a.e += ex - 1023;
  a.m *= __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(x));
}
__device__ __forceinline__ void prod_renorm(ProdAcc &a) {
  int hi = __double2hiint(a.m);
  int ex = (hi >> 20) & 0x7ff;
  a.e += ex - 1023;
  a.m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(a.m));
}
__device__ __forceinline__ void prod_merge(ProdAcc &a, double m, int e) { a.m *= m; a.e += e; }
__device__ __forceinline__ double prod_log10(const ProdAcc &a) { return log10(a.m) + (double)a.e * kLog10_2; }

// ---- Brent as a resumable state machine (Gold:81-177 entered the way NucFam:432-444 enters it) -----
// fa and fc of OptimizeFrequency are evaluated by the reference but never read by Brent, so they are
// not evaluated here.  Usage:  brent_begin(s) -> evaluate f(s.u) -> while (brent_feed(s, f, tol)) evaluate f(s.u).
struct BrentState {
  double a, c, min, w, v, fmin, fw, fv, delta, d, u;
  int iter;
  int first;
};
#define PM_ITMAX 200
#define PM_ZEPS 3.0e-10
#define PM_CGOLD 0.38196601

__device__ __forceinline__ double sign_d(double a, double b) { return b >= 0 ? fabs(a) : -fabs(a); }

__device__ inline void brent_begin(BrentState &s) {
  s.a = 0.0001; s.c = 0.5;  // a < c already (Gold:85-89)
  s.min = s.w = s.v = 0.9999;
  s.delta = 0.0; s.d = 0.0;
  s.u = 0.9999;
  s.iter = 0; s.first = 1;
}
// Consumes fu = f(s.u).  Returns true if another evaluation (at the new s.u) is needed.
__device__ inline bool brent_feed(BrentState &s, double fu, double tol) {
  if (s.first) {
    s.fmin = s.fw = s.fv = fu;
    s.first = 0;
  } else {
    double u = s.u;
    if (fu <= s.fmin) {
      if (u >= s.min) s.a = s.min; else s.c = s.min;
      s.v = s.w; s.w = s.min; s.min = u;
      s.fv = s.fw; s.fw = s.fmin; s.fmin = fu;
    } else {
      if (u < s.min) s.a = u; else s.c = u;
      if (fu <= s.fw || s.w == s.min) {
        s.v = s.w; s.w = u;
        s.fv = s.fw; s.fw = fu;
      } else if (fu <= s.fv || s.v == s.min || s.v == s.w) {
        s.v = u; s.fv = fu;
      }
    }
  }
  s.iter++;
  if (s.iter > PM_ITMAX) return false;  // "Brent got stuck": the reference warns and returns fmin
  double middle = 0.5 * (s.a + s.c);
  double tol1 = tol * fabs(s.min) + PM_ZEPS;
  double tol2 = 2.0 * tol1;
  if (fabs(s.min - middle) <= (tol2 - 0.5 * (s.c - s.a))) return false;
  if (fabs(s.delta) > tol1) {
    double r = (s.min - s.w) * (s.fmin - s.fv);
    double q = (s.min - s.v) * (s.fmin - s.fw);
    double p = (s.min - s.v) * q - (s.min - s.w) * r;
    q = 2.0 * (q - r);
    if (q > 0.0) p = -p;
    q = fabs(q);
    double temp = s.delta;
    s.delta = s.d;
    if (fabs(p) >= fabs(0.5 * q * temp) || p <= q * (s.a - s.min) || p >= q * (s.c - s.min)) {
      s.delta = s.min >= middle ? s.a - s.min : s.c - s.min;
      s.d = PM_CGOLD * s.delta;
    } else {
      s.d = p / q;
      double u = s.min + s.d;
      if (u - s.a < tol2 || s.c - u < tol2) s.d = sign_d(tol1, middle - s.min);
    }
  } else {
    s.delta = s.min >= middle ? s.a - s.min : s.c - s.min;
    s.d = PM_CGOLD * s.delta;
  }
  s.u = fabs(s.d) >= tol1 ? s.min + s.d : s.min + sign_d(tol1, s.d);
  return true;
}

// ---- hypothesis bookkeeping (main:439-553, NucFam:1693-1749) -------------------------------------
__device__ __forceinline__ void hyp_alleles(int h, int ref, int &a1, int &a2) {
  int ts = poly_ts(ref), t1 = poly_tvs1(ref), t2 = poly_tvs2(ref);
  switch (h) {
    case 0: a1 = ref; a2 = (ref == 4) ? ref - 1 : ref + 1; break;  // main:458
    case 1: a1 = ref; a2 = ts; break;
    case 2: a1 = ref; a2 = t1; break;
    case 3: a1 = ref; a2 = t2; break;
    case 4: a1 = ts; a2 = t1; break;
    case 5: a1 = ts; a2 = t2; break;
    default: a1 = t1; a2 = t2; break;
  }
}

// CalcVarPosterior for the first n hypotheses; fills maxidx, varPostProb, polyQual and the alleles
// famlk[0] holds afterwards (for a mono winner: ref + best alternative among H1..H3, NucFam:1664-1683).
__device__ inline void var_posterior(pm_site_result &r, int ref, int n) {
  int maxidx = 0;
  double mx = r.varllk[0];
  for (int i = 0; i < n; i++) if (mx < r.varllk[i]) { mx = r.varllk[i]; maxidx = i; }
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += exp10(r.varllk[i] - r.varllk[maxidx]);
  r.var_post_prob = 1 / sum;
  int a1, a2;
  if (maxidx == 0) {
    int idx = 1;
    double m2 = r.varllk[1];
    for (int i = 1; i < 4; i++) if (m2 < r.varllk[i]) { m2 = r.varllk[i]; idx = i; }
    hyp_alleles(idx, ref, a1, a2);
  } else {
    hyp_alleles(maxidx, ref, a1, a2);
  }
  r.allele1 = (uint8_t)a1; r.allele2 = (uint8_t)a2;
  r.maxidx = (int8_t)maxidx;
  r.n_hyp = (uint8_t)n;
  r.poly_qual = (r.var_post_prob > 0.9999999999) ? 100.0 : -10 * log10(1 - r.var_post_prob);
}

// VCF mode (src/PedVCF.cpp:140-156): the site prior is dropped from llk_alt by operator precedence
// (`log10(polyPrior * isTs(...) ? ts : tv)`), isTs is true only for A>G and C>T, indels use the SNP prior.
__device__ inline void vcf_record_result(const DevRun *run, pm_site_result &r, int a1, int a2, bool indel, double mono, double poly, double freq) {
  const bool is_ts = (a1 == 1 && a2 == 3) || (a1 == 2 && a2 == 4);
  const double llk_alt = (indel ? run->vcf_log_indel : (is_ts ? run->vcf_log_ts : run->vcf_log_tv)) + poly;
  const double llk_ref = run->log_1m_prior + mono;
  double qual;
  if (llk_alt - llk_ref > 10) qual = 10.0 * (llk_alt - llk_ref);
  else {
    const double posterior = 1 / (1 + pow(10.0, llk_ref - llk_alt));
    qual = -10 * log10(1 - posterior);
    r.var_post_prob = posterior;
  }
  r.varllk[0] = llk_ref; r.varllk[1] = llk_alt;
  r.varllk_noprior[0] = mono; r.varllk_noprior[1] = poly;
  r.varfreq[0] = 1.0; r.varfreq[1] = freq;
  r.poly_qual = qual;
  r.freq = freq;
  r.allele1 = (uint8_t)a1; r.allele2 = (uint8_t)a2;
  r.maxidx = 1; r.n_hyp = 2;
  r.status = PM_SITE_EMITTED;
}

__device__ __forceinline__ int best3(double p11, double p12, double p22) {  // NucFam:1564-1571
  int b = 0; double best = p11;
  if (p12 > best) { best = p12; b = 1; }
  if (p22 > best) { best = p22; b = 2; }
  return b;
}
__device__ __forceinline__ uint8_t gq_of(double pb) {  // NucFam:1819-1820
  int q;
  if (pb > 0.9999999999) q = 100;
  else q = (int)(-10. * log10(1. - pb) + 0.5);
  return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

}  // namespace pm
