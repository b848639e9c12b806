// pm_es.cuh — thread-serial Elston-Stewart peel and the shared-memory tables, used by the narrow site kernel
// (pm_kernels.cu) and by the posterior kernel (pm_post.cu).
#pragma once
#include "pm_device.cuh"

namespace pm {

// ================================================================================================
// Elston–Stewart peel, thread-serial (ES:990-1057, 1078-1395).  A = 3 (bi-allelic) or 10 (--denovo).
// pin_person >= 0 zeroes that person's penetrance except genotype pin_geno (FillZeroPenetrance,
// FLSeq:327-356), which is how the reference gets per-person posteriors in extended pedigrees.
// ================================================================================================
__device__ __forceinline__ double tba(int i, int j, int k) {  // transmission_BA, ES:824-832
  double ti = 0.5 * i, tj = 0.5 * j;
  return k == 0 ? (1 - ti) * (1 - tj) : (k == 2 ? ti * tj : ti * (1 - tj) + (1 - ti) * tj);
}
// GetTransmissionProb_BA (ES:1059-1075) with the chrX-to-female / chrX-to-male / chrY / mitochondrial tables
// (ES:834-924) in closed form; `sex` is the offspring's.
__device__ __forceinline__ double tba_cls(int cls, int sex, int i, int j, int k) {
  if (cls == PM_CHR_AUTO) return tba(i, j, k);
  const bool male = sex == 1;
  if (cls == PM_CHR_Y && !male) return 1.0;
  if (i == 1) return 0.0;  // a heterozygous father does not exist on these chromosomes
  if (cls == PM_CHR_X) {
    if (!male) return tba(i, j, k);
    const double tj = 0.5 * j;  // a son gets his only copy from the mother
    return k == 0 ? 1 - tj : (k == 2 ? tj : 0.0);
  }
  if (cls == PM_CHR_Y) return (k == i) ? 1.0 : 0.0;      // the father's copy
  if (j == 1) return 0.0;                                // MT: the mother's copy
  return (k == j) ? 1.0 : 0.0;
}

// ================================================================================================
// Ten-state peel (--denovo), factorised.  The reference contracts dense 10 x 10 x 10 tensors
// (transmission / transmission_denovo, ES:752-810) in every step: ~36 kflop per likelihood of the 20-member CEPH
// pedigree.  The tensors have structure the contraction can use:
//   * transmission[i][j][k] = 1/4 for each of the four (allele of i, allele of j) pairs that make genotype k, and
//     transmission_denovo = transmission x M (M = the 10 x 10 genotype mutation matrix).  So
//       sum_k tden[i][j][k] pc[k] = 1/4 sum_{x in i, y in j} (M pc)[g(x,y)]
//     = 1/4 (R[a_i][j] + R[b_i][j]),  R[x][j] = A[x][c_j] + A[x][d_j],  A[x][y] = (M pc)[g(x,y)]:
//     100 FMAs for M pc, 40 + 55 additions, instead of 1,000 FMAs per child;
//   * the result is symmetric in (i, j), so a marriage partial is kept as 55 numbers (i <= j);
//   * a couple without marriage partial sends to its child  sum_ij pf[i] pm[j] T[i][j][k] = the gamete product
//     gf[x] gm[y] (+ gf[y] gm[x]), gf[x] = sum_i pf[i] * (copies of x in i) / 2: ~60 flops instead of 3,000.
// Same sums in a different order: results agree with the dense contraction to ~1e-15 relative; a likelihood that
// is exactly zero (no compatible genotype configuration) is exactly zero here too (all terms are non-negative).
// Quirk kept: a couple WITH a marriage partial sends through `transmission` without mutation (ES:1391).
// ================================================================================================
__device__ __forceinline__ constexpr int g10(int a, int b) { return a <= b ? a * 4 - a * (a - 1) / 2 + (b - a) : b * 4 - b * (b - 1) / 2 + (a - b); }
__device__ __forceinline__ constexpr int sym55(int i, int j) { return i <= j ? i * 10 - i * (i - 1) / 2 + (j - i) : j * 10 - j * (j - 1) / 2 + (i - j); }
__device__ __forceinline__ constexpr int al10(int g, int w) {
  // alleles of genotype g in the order AA AC AG AT CC CG CT GG GT TT
  return w == 0 ? (g < 4 ? 0 : (g < 7 ? 1 : (g < 9 ? 2 : 3))) : (g < 4 ? g : (g < 7 ? g - 3 : (g < 9 ? g - 5 : 3)));
}

template <bool NA, typename RecPtr>
__device__ double es_likelihood10(const DevRun *__restrict__ run, const DevFam f, RecPtr recs, int g11, int g12, int g22, bool denovo, double freq,
                                  const double *__restrict__ lut, const double *__restrict__ mut, int pin_person, int pin_geno, int cls_) {
  double part[kMaxEsPersons * 10];
  double mp[kMaxMp * 55];
  const double q = 1.0 - freq;
  const int cls = NA ? cls_ : PM_CHR_AUTO;
  const uint8_t *sexes = NA ? run->sex + f.first : nullptr;
  for (int i = 0; i < f.size; i++) {  // SetFounderPriors / InitializePartials, ES:643-664, 1434-1446
    const uint4 r = recs[f.first + i];
    double pr[3] = {freq * freq, 2 * freq * q, q * q};
    if constexpr (NA) {
      const bool male = sexes[i] == 1;
      if (cls == PM_CHR_MT || ((cls == PM_CHR_X || cls == PM_CHR_Y) && male)) { pr[0] = freq; pr[1] = 0.0; pr[2] = q; }
      else if (cls == PM_CHR_Y) { pr[0] = pr[1] = pr[2] = 1.0; }
    }
#pragma unroll
    for (int g = 0; g < 10; g++) {
      double pen = lut[rec_lk(r, g)];
      if (i == pin_person && g != pin_geno) pen = 0.0;
      if (i < f.founders) pen *= g == g11 ? pr[0] : (g == g12 ? pr[1] : (g == g22 ? pr[2] : 0.0));
      part[i * 10 + g] = pen;
    }
  }
  const DevStep *steps = run->steps + f.step_first;
  for (int s = 0; s < f.n_steps; s++) {
    const DevStep st = steps[s];
    if (st.type == PM_PEEL_CHILD_TO_PARENTS) {
      // A run of children of the same couple (consecutive steps into one marriage partial: 14 of them in the CEPH
      // pedigree): the 55 products are accumulated in REGISTERS over the whole run and stored once — the per-child
      // read-modify-write of the marriage partial in local memory was what bound this kernel (ncu: the top stall was
      // the local-memory scoreboard, 31 KB of DRAM write-back per site).
      double acc[55];
      bool have = false;
      if (!st.flag) {
#pragma unroll
        for (int e = 0; e < 55; e++) acc[e] = mp[st.mp * 55 + e];
        have = true;
      }
      int s2 = s;
      for (; s2 < f.n_steps; s2++) {
        const DevStep sk = steps[s2];
        if (sk.type != PM_PEEL_CHILD_TO_PARENTS || sk.mp != st.mp) break;
        double pc[10], pcm[10];
#pragma unroll
        for (int k = 0; k < 10; k++) pc[k] = part[sk.from0 * 10 + k];
        if (denovo) {
#pragma unroll
          for (int m = 0; m < 10; m++) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int k = 0; k < 5; k++) { a = fma(mut[m * 10 + k], pc[k], a); b = fma(mut[m * 10 + 5 + k], pc[5 + k], b); }
            pcm[m] = 0.25 * (a + b);
          }
        } else {
#pragma unroll
          for (int m = 0; m < 10; m++) pcm[m] = 0.25 * pc[m];
        }
#pragma unroll
        for (int i = 0; i < 10; i++)
#pragma unroll
          for (int j = i; j < 10; j++) {
            // 1/4 sum over (allele of i, allele of j) of (M pc)[g(x, y)]
            const double v = (pcm[g10(al10(i, 0), al10(j, 0))] + pcm[g10(al10(i, 0), al10(j, 1))]) +
                             (pcm[g10(al10(i, 1), al10(j, 0))] + pcm[g10(al10(i, 1), al10(j, 1))]);
            acc[sym55(i, j)] = have ? acc[sym55(i, j)] * v : v;  // a fresh marriage partial starts at 1
          }
        have = true;
      }
#pragma unroll
      for (int e = 0; e < 55; e++) mp[st.mp * 55 + e] = acc[e];
      s = s2 - 1;
    } else if (st.type == PM_PEEL_SPOUSE_TO_SPOUSE) {
      const double *pf = part + st.from0 * 10;
      double *pt = part + st.to0 * 10;
      if (st.mp < 0) {
        double sum = 0.0;
        for (int j = 0; j < 10; j++) sum += pf[j];
        for (int i = 0; i < 10; i++) pt[i] *= sum;
      } else {
        const double *m = mp + st.mp * 55;  // symmetric: the same whichever spouse is folded in
        double ps[10];
#pragma unroll
        for (int j = 0; j < 10; j++) ps[j] = pf[j];
#pragma unroll
        for (int i = 0; i < 10; i++) {
          double sum = 0.0;
#pragma unroll
          for (int j = 0; j < 10; j++) sum = fma(ps[j], m[sym55(i, j)], sum);
          pt[i] *= sum;
        }
      }
    } else {
      const double *pf = part + st.from0 * 10, *pm_ = part + st.from1 * 10;
      double *pc = part + st.to0 * 10;
      double w[10];
      if (st.mp < 0) {  // gamete vectors
        double gf[4] = {0, 0, 0, 0}, gm[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 10; i++) {
          const double hf = 0.5 * pf[i], hm = 0.5 * pm_[i];
          gf[al10(i, 0)] += hf; gf[al10(i, 1)] += hf;
          gm[al10(i, 0)] += hm; gm[al10(i, 1)] += hm;
        }
#pragma unroll
        for (int k = 0; k < 10; k++) {
          const int x = al10(k, 0), y = al10(k, 1);
          w[k] = x == y ? gf[x] * gm[x] : fma(gf[x], gm[y], gf[y] * gm[x]);
        }
        if (denovo) {  // ES:1383: transmission_denovo = transmission x M
          double o[10];
#pragma unroll
          for (int k = 0; k < 10; k++) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 10; m++) sum = fma(w[m], mut[m * 10 + k], sum);
            o[k] = sum;
          }
#pragma unroll
          for (int k = 0; k < 10; k++) w[k] = o[k];
        }
      } else {  // ES:1391: with a marriage partial the plain transmission tensor, also under --denovo
        const double *m = mp + st.mp * 55;
#pragma unroll
        for (int k = 0; k < 10; k++) w[k] = 0.0;
#pragma unroll
        for (int i = 0; i < 10; i++)
#pragma unroll
          for (int j = 0; j < 10; j++) {
            const double v = 0.25 * (pf[i] * m[sym55(i, j)] * pm_[j]);
            w[g10(al10(i, 0), al10(j, 0))] += v; w[g10(al10(i, 0), al10(j, 1))] += v;
            w[g10(al10(i, 1), al10(j, 0))] += v; w[g10(al10(i, 1), al10(j, 1))] += v;
          }
      }
#pragma unroll
      for (int k = 0; k < 10; k++) pc[k] *= w[k];
    }
  }
  const double *pfin = part + steps[f.n_steps - 1].to0 * 10;
  double lk = 0.0;
  for (int i = 0; i < 10; i++) lk += pfin[i];
  return lk;
}

// ================================================================================================
// Three-state peel (bi-allelic).  Arithmetic as the reference's (ES:990-1057 with the _BA tables), but a person's partial
// lives in local memory only from the moment a peel step multiplies into it: until then it is its initial value, three
// table look-ups on the person's record, recomputed where it is read (a leaf child is read once).  A run of children of
// one couple accumulates the nine products of the marriage partial in registers and stores them once.  For the 20-member
// CEPH pedigree that is 2 persons and one marriage partial in local memory instead of 20 and 14 read-modify-write rounds
// (ncu of the round-2 kernel: long-scoreboard stall 10 warps per issue cycle, 6 KB of local-memory write-back to DRAM
// per site).
// NA = false is the autosomal code with none of the chrX / chrY / MT rules compiled in (cls_ ignored): the rules sit in
// the innermost loops of the peel and cost the narrow kernel 40 % when they were runtime branches.
// ================================================================================================
template <bool NA, typename RecPtr>
__device__ double es_likelihood3(const DevRun *__restrict__ run, const DevFam f, RecPtr recs, int g11, int g12, int g22, double freq,
                                 const double *__restrict__ lut, int pin_person, int pin_geno, int cls_) {
  constexpr int A = 3;
  double part[kMaxEsPersons * A];
  double mp[kMaxMp * A * A];
  unsigned long long live = 0ull;  // bit i: person i's partial has been materialised in part[]
  static_assert(kMaxEsPersons <= 64, "one bit per family member");
  const int gi[3] = {g11, g12, g22};
  const double q = 1.0 - freq;
  const int cls = NA ? cls_ : PM_CHR_AUTO;
  const uint8_t *sexes = NA ? run->sex + f.first : nullptr;
  // SetFounderPriors_BA + InitializePartials_BA, ES:643-687, 1449-1465
  auto initial = [&](int i, double *o) {
    const uint4 r = recs[f.first + i];
    double pr[3] = {freq * freq, 2 * freq * q, q * q};
    bool yfemale = false;
    if constexpr (NA) {
      const bool male = sexes[i] == 1;
      if (cls == PM_CHR_MT || ((cls == PM_CHR_X || cls == PM_CHR_Y) && male)) { pr[0] = freq; pr[1] = 0.0; pr[2] = q; }
      else if (cls == PM_CHR_Y) { pr[0] = pr[1] = pr[2] = 1.0; }
      yfemale = cls == PM_CHR_Y && sexes[i] == 2;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double pen = lut[rec_lk(r, gi[j])];
      if (i == pin_person && gi[j] != pin_geno) pen = 0.0;
      o[j] = yfemale ? 1.0 : ((i < f.founders) ? pr[j] * pen : pen);
    }
  };
  auto load = [&](int i, double *o) {
    if ((live >> i) & 1ull) { o[0] = part[i * A]; o[1] = part[i * A + 1]; o[2] = part[i * A + 2]; }
    else initial(i, o);
  };
  const DevStep *steps = run->steps + f.step_first;
  for (int s = 0; s < f.n_steps; s++) {
    const DevStep st = steps[s];
    if (st.type == PM_PEEL_CHILD_TO_PARENTS) {
      double acc[A * A];
      if (!st.flag) {
#pragma unroll
        for (int e = 0; e < A * A; e++) acc[e] = mp[st.mp * A * A + e];
      }
      int s2 = s;
      for (; s2 < f.n_steps; s2++) {
        const DevStep sk = steps[s2];
        if (sk.type != PM_PEEL_CHILD_TO_PARENTS || sk.mp != st.mp) break;
        double pc[A];
        load(sk.from0, pc);
#pragma unroll
        for (int i = 0; i < A; i++)
#pragma unroll
          for (int j = 0; j < A; j++) {
            double sum = 0;
            if constexpr (NA) { for (int k = 0; k < 3; k++) sum += tba_cls(cls, sexes[sk.from0], i, j, k) * pc[k]; }
            else { for (int k = 0; k < 3; k++) sum += tba(i, j, k) * pc[k]; }
            acc[i * A + j] = sk.flag ? sum : acc[i * A + j] * sum;  // a fresh marriage partial starts at 1
          }
      }
#pragma unroll
      for (int e = 0; e < A * A; e++) mp[st.mp * A * A + e] = acc[e];
      s = s2 - 1;
    } else if (st.type == PM_PEEL_SPOUSE_TO_SPOUSE) {
      double pf[A], pt[A];
      load(st.from0, pf);
      load(st.to0, pt);
      if (st.mp < 0) {
        double sum = 0.0;
        for (int j = 0; j < A; j++) sum += pf[j];
        for (int i = 0; i < A; i++) pt[i] *= sum;
      } else {
        const double *m = mp + st.mp * A * A;
        for (int i = 0; i < A; i++) {
          double sum = 0.0;
          if (st.flag) for (int j = 0; j < A; j++) sum += pf[j] * m[j * A + i];
          else for (int j = 0; j < A; j++) sum += pf[j] * m[i * A + j];
          pt[i] *= sum;
        }
      }
      part[st.to0 * A] = pt[0]; part[st.to0 * A + 1] = pt[1]; part[st.to0 * A + 2] = pt[2];
      live |= 1ull << st.to0;
    } else {
      double pf[A], pm_[A], pc[A];
      load(st.from0, pf);
      load(st.from1, pm_);
      load(st.to0, pc);
      const double *m = st.mp >= 0 ? mp + st.mp * A * A : nullptr;
      for (int k = 0; k < A; k++) {
        double sum = 0.0;
        for (int i = 0; i < A; i++)
          for (int j = 0; j < A; j++) {
            double t;
            if constexpr (NA) t = tba_cls(cls, sexes[st.to0], i, j, k); else t = tba(i, j, k);
            if (m) sum += pf[i] * m[i * A + j] * pm_[j] * t;
            else sum += pf[i] * pm_[j] * t;
          }
        pc[k] *= sum;
      }
      part[st.to0 * A] = pc[0]; part[st.to0 * A + 1] = pc[1]; part[st.to0 * A + 2] = pc[2];
      live |= 1ull << st.to0;
    }
  }
  double pfin[A];
  load(steps[f.n_steps - 1].to0, pfin);
  double lk = 0.0;
  for (int i = 0; i < A; i++) lk += pfin[i];
  return lk;
}

// The three-state peel as it was before the lazy partials (every partial initialised up front, marriage partials updated in
// place): smaller code.  The --denovo instances of the thread-per-site kernel use it for their rare bi-allelic refit, where
// the register-hungry form above, inlined next to the ten-state peel (or called out of line), cost the CEPH --denovo
// instance 9 % (33.6-33.9 vs 37.0 M sites/s).
template <bool NA, typename RecPtr>
__device__ double es_likelihood3_eager(const DevRun *__restrict__ run, const DevFam f, RecPtr recs, int g11, int g12, int g22,
                                     bool denovo, double freq, const double *__restrict__ lut,
                                     const double *__restrict__ mut, int pin_person, int pin_geno, int cls_) {
  constexpr int A = 3;
  double part[kMaxEsPersons * A];
  double mp[kMaxMp * A * A];
  const int gi[3] = {g11, g12, g22};
  const double q = 1.0 - freq;
  const int cls = NA ? cls_ : PM_CHR_AUTO;
  const uint8_t *sexes = NA ? run->sex + f.first : nullptr;
  for (int i = 0; i < f.size; i++) {
    uint4 r = recs[f.first + i];
    double pr[3] = {freq * freq, 2 * freq * q, q * q};  // SetFounderPriors{,_BA}, ES:643-687
    if constexpr (NA) {
      const bool male = sexes[i] == 1;
      if (cls == PM_CHR_MT || ((cls == PM_CHR_X || cls == PM_CHR_Y) && male)) { pr[0] = freq; pr[1] = 0.0; pr[2] = q; }
      else if (cls == PM_CHR_Y) { pr[0] = pr[1] = pr[2] = 1.0; }
    }
    if (A == 3) {
      bool yfemale = false;
      if constexpr (NA) yfemale = cls == PM_CHR_Y && sexes[i] == 2;
      for (int j = 0; j < 3; j++) {
        double pen = lut[rec_lk(r, gi[j])];
        if (i == pin_person && gi[j] != pin_geno) pen = 0.0;
        part[i * 3 + j] = yfemale ? 1.0 : ((i < f.founders) ? pr[j] * pen : pen);  // InitializePartials_BA, ES:1449-1465
      }
    } else {
      for (int g = 0; g < 10; g++) {
        double pen = lut[rec_lk(r, g)];
        if (i == pin_person && g != pin_geno) pen = 0.0;
        if (i < f.founders) {  // InitializePartials, ES:1434-1446
          double prior = g == g11 ? pr[0] : (g == g12 ? pr[1] : (g == g22 ? pr[2] : 0.0));
          part[i * 10 + g] = prior * pen;
        } else {
          part[i * 10 + g] = pen;
        }
      }
    }
  }
  const DevStep *steps = run->steps + f.step_first;
  for (int s = 0; s < f.n_steps; s++) {
    const DevStep st = steps[s];
    if (st.type == PM_PEEL_CHILD_TO_PARENTS) {
      double *m = mp + st.mp * A * A;
      const double *pc = part + st.from0 * A;
      for (int i = 0; i < A; i++)
        for (int j = 0; j < A; j++) {
          double sum = 0;
          if (A == 3) {
            if constexpr (NA) { for (int k = 0; k < 3; k++) sum += tba_cls(cls, sexes[st.from0], i, j, k) * pc[k]; }
            else { for (int k = 0; k < 3; k++) sum += tba(i, j, k) * pc[k]; }
          }
          m[i * A + j] = st.flag ? sum : m[i * A + j] * sum;  // a fresh marriage partial starts at 1
        }
    } else if (st.type == PM_PEEL_SPOUSE_TO_SPOUSE) {
      const double *pf = part + st.from0 * A;
      double *pt = part + st.to0 * A;
      if (st.mp < 0) {
        double sum = 0.0;
        for (int j = 0; j < A; j++) sum += pf[j];
        for (int i = 0; i < A; i++) pt[i] *= sum;
      } else {
        const double *m = mp + st.mp * A * A;
        for (int i = 0; i < A; i++) {
          double sum = 0.0;
          if (st.flag) for (int j = 0; j < A; j++) sum += pf[j] * m[j * A + i];
          else for (int j = 0; j < A; j++) sum += pf[j] * m[i * A + j];
          pt[i] *= sum;
        }
      }
    } else {
      const double *pf = part + st.from0 * A, *pm_ = part + st.from1 * A;
      double *pc = part + st.to0 * A;
      const double *m = st.mp >= 0 ? mp + st.mp * A * A : nullptr;
      for (int k = 0; k < A; k++) {
        double sum = 0.0;
        for (int i = 0; i < A; i++)
          for (int j = 0; j < A; j++) {
            double t;
            if constexpr (NA) t = tba_cls(cls, sexes[st.to0], i, j, k); else t = tba(i, j, k);
            if (m) sum += pf[i] * m[i * A + j] * pm_[j] * t;
            else sum += pf[i] * pm_[j] * t;
          }
        pc[k] *= sum;
      }
    }
  }
  const double *pfin = part + steps[f.n_steps - 1].to0 * A;
  double lk = 0.0;
  for (int i = 0; i < A; i++) lk += pfin[i];
  return lk;
}

template <int A, bool NA, typename RecPtr>
__device__ double es_likelihood_impl(const DevRun *__restrict__ run, const DevFam f, RecPtr recs, int g11, int g12, int g22,
                                     bool denovo, double freq, const double *__restrict__ lut,
                                     const double *__restrict__ mut, int pin_person, int pin_geno, int cls_) {
  static_assert(A == 3 || A == 10, "three-state (bi-allelic) or ten-state (--denovo) peel");
  if constexpr (A == 10) return es_likelihood10<NA>(run, f, recs, g11, g12, g22, denovo, freq, lut, mut, pin_person, pin_geno, cls_);
  else return es_likelihood3<NA>(run, f, recs, g11, g12, g22, freq, lut, pin_person, pin_geno, cls_);
}

template <int A, typename RecPtr>
__device__ __forceinline__ double es_likelihood(const DevRun *__restrict__ run, const DevFam f, RecPtr recs, int g11, int g12, int g22,
                                                bool denovo, double freq, const double *__restrict__ lut,
                                                const double *__restrict__ mut, int pin_person,
                                                int pin_geno, int cls = PM_CHR_AUTO) {
  if (cls == PM_CHR_AUTO) return es_likelihood_impl<A, false>(run, f, recs, g11, g12, g22, denovo, freq, lut, mut, pin_person, pin_geno, cls);
  return es_likelihood_impl<A, true>(run, f, recs, g11, g12, g22, denovo, freq, lut, mut, pin_person, pin_geno, cls);
}

// shared tables at the start of dynamic shared memory
struct SmemTables {
  double lut[256];
  double mut[100];
};

__device__ __forceinline__ void load_tables(const DevRun *run, SmemTables *t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) t->lut[i] = run->lut[i];
  for (int i = threadIdx.x; i < 100; i += blockDim.x) t->mut[i] = run->mut[i];
}

}  // namespace pm
