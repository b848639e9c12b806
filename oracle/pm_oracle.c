/*
 * pm_oracle.c — CPU restatement of polymutt 0.13's per-site family-likelihood path (GLF input).
 *
 * TEST INFRASTRUCTURE ONLY (see pm_oracle.h).  Plain C99, double precision, the reference's own
 * operation order wherever the order can change a bit.  Every function cites the reference
 * file:line it follows ("NucFam" = src/NucFamGenotypeLikelihood.cpp, "FLSeq" =
 * src/FamilyLikelihoodSeq.cpp, "ES" = src/FamilyLikelihoodES.cpp, "Gold" = core/MathGold.cpp,
 * "main" = src/main.cpp).  Parity status: pinned (see pm_oracle.h).
 *
 * Autosomes only: chrX / chrY / MT sites return PM_EUNSUPPORTED.
 */
#define _GNU_SOURCE
#include "pm_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define pow10(x) exp10(x) /* the reference is built with the same mapping (oracle/build_ref.sh) */

static char g_err[512];
static void set_err(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char *pmo_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------------ */
/* constants: core/MathConstant.h:14-25 */
#define ITMAX 200
#define ZEPS 3.0e-10
#define CGOLD 0.38196601
static double sign_d(double a, double b) { return b >= 0 ? fabs(a) : -fabs(a); }

/* core/glfHandler.h:102-106 */
static int GenotypeIndex(int base1, int base2) {
  return base1 < base2 ? (base1 - 1) * (10 - base1) / 2 + (base2 - base1)
                       : (base2 - 1) * (10 - base2) / 2 + (base1 - base2);
}
/* src/PedigreeGLF.h:15-53 */
static int poly_ts(int r) { static const int t[5] = {0, 3, 4, 1, 2}; return t[r]; }
static int poly_tvs1(int r) { static const int t[5] = {0, 2, 1, 2, 1}; return t[r]; }
static int poly_tvs2(int r) { static const int t[5] = {0, 4, 3, 4, 3}; return t[r]; }

/* core/BaseQualityHelper.cpp:12-13 */
void pmo_fill_lut(double *lut) {
  for (int i = 0; i <= 255; i++) lut[i] = pow(0.1, i * 0.1);
}

/* src/MutationModel.cpp:15-30 and 46-90 */
void pmo_genotype_mutation_matrix(double mu, double tstv, double *m100) {
  double aM[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (i == j) aM[i][j] = 1 - mu;
      else aM[i][j] = (1 - mu) / 3;
  if (tstv != 0.0) {
    aM[0][2] = aM[2][0] = aM[1][3] = aM[3][1] = mu / 3 * (3 - 3 / (1 + tstv));
    aM[0][1] = aM[0][3] = aM[1][0] = aM[1][2] = aM[2][1] = aM[2][3] = aM[3][0] = aM[3][2] =
        mu / 3 * (0.5 / (1 + tstv) * 3);
  }
  double mutRate16[16][16];
  int fromIdx = -1, toIdx;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      fromIdx++;
      toIdx = -1;
      for (int ii = 0; ii < 4; ii++)
        for (int jj = 0; jj < 4; jj++) {
          toIdx++;
          mutRate16[fromIdx][toIdx] = (aM[i][ii] * aM[j][jj]);
        }
    }
  static const int hetOrdered1[6] = {2, 3, 4, 7, 8, 12};
  static const int hetOrdered2[6] = {5, 9, 13, 10, 14, 15};
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 16; j++) mutRate16[j][hetOrdered1[i] - 1] += mutRate16[j][hetOrdered2[i] - 1];
  static const int unOrderedIdx[10] = {1, 2, 3, 4, 6, 7, 8, 11, 12, 16};
  for (int i = 0; i < 10; i++)
    for (int j = 0; j < 10; j++) m100[i * 10 + j] = mutRate16[unOrderedIdx[i] - 1][unOrderedIdx[j] - 1];
}

/* ------------------------------------------------------------------------------------------ */
/* ES_Peeling (ES:46-277, 454-512).  IntArray restated as a tiny vector with the reference's
 * Delete semantics (core/IntArray.cpp:81-87: the slot past the end keeps its stale value). */
typedef struct { int n; int v[64]; } ivec;
static void iv_push(ivec *a, int x) { a->v[a->n++] = x; }
static void iv_del(ivec *a, int idx) {
  a->n--;
  if (a->n - idx) memmove(a->v + idx, a->v + idx + 1, sizeof(int) * (size_t)(a->n - idx));
}
static int iv_find(const ivec *a, int x) {
  for (int i = 0; i < a->n; i++) if (a->v[i] == x) return i;
  return -1;
}
typedef struct {
  int n;
  ivec *parents, *offspring, *spouses;
  const uint8_t *sex;
} peel_t;
static int es_isFinal(peel_t *p, int i) { return p->parents[i].n == 0 && p->spouses[i].n == 0 && p->offspring[i].n == 0; }
static int es_isLeaf(peel_t *p, int i) { return p->offspring[i].n == 0 && p->spouses[i].n == 0; }
static int es_isPeripheral(peel_t *p, int i) { return p->offspring[i].n == 0 && p->parents[i].n == 0 && p->spouses[i].n == 1; }
static int es_isRoof(peel_t *p, int i) {
  if (p->spouses[i].n != 1) return 0;
  int s = p->spouses[i].v[0];
  return p->spouses[s].n == 1 && p->parents[i].n == 0 && p->parents[s].n == 0 && p->offspring[i].n == 1 &&
         p->offspring[s].n == 1;
}
typedef struct { int n; int a[64], b[64]; } pairvec;
static int pv_find(const pairvec *v, int f, int s) { /* ES:472-483 find_element: either order */
  for (int i = 0; i < v->n; i++)
    if ((v->a[i] == f && v->b[i] == s) || (v->a[i] == s && v->b[i] == f)) return i;
  return -1;
}
static void pv_erase(pairvec *v, int idx) {
  for (int i = idx; i + 1 < v->n; i++) { v->a[i] = v->a[i + 1]; v->b[i] = v->b[i + 1]; }
  v->n--;
}
static void es_UpdateRoof(peel_t *p, pairvec *roof, int index) { /* ES:461-470 */
  int s = p->spouses[index].v[0];
  if (pv_find(roof, index, s) >= 0) return;
  roof->a[roof->n] = index; roof->b[roof->n] = s; roof->n++;
}

int pmo_build_peel_order(int32_t n, const int32_t *father, const int32_t *mother, const uint8_t *sex,
                         pm_peel_step *steps) {
  if (n > 60) { set_err("oracle peel builder supports families up to 60 members"); return PM_EINVAL; }
  peel_t P;
  P.n = n; P.sex = sex;
  P.parents = calloc((size_t)n, sizeof(ivec));
  P.offspring = calloc((size_t)n, sizeof(ivec));
  P.spouses = calloc((size_t)n, sizeof(ivec));
  int rc = 0, nsteps = 0;
  /* SetupConnections, ES:46-78 */
  {
    int cf[64 * 64];
    memset(cf, 0, sizeof cf);
    for (int i = 0; i < n; i++) {
      if (father[i] < 0 || mother[i] < 0) continue;
      int fa = father[i], mo = mother[i];
      iv_push(&P.parents[i], fa); iv_push(&P.parents[i], mo);
      iv_push(&P.offspring[fa], i); iv_push(&P.offspring[mo], i);
      if (cf[fa * 64 + mo] == 0) { iv_push(&P.spouses[fa], mo); iv_push(&P.spouses[mo], fa); }
      cf[fa * 64 + mo]++;
    }
  }
  ivec leaf = {0}, peripheral = {0};
  pairvec roof; roof.n = 0;
  /* BuildInitialPeelable, ES:80-115 */
  {
    int visited[64]; memset(visited, 0, sizeof visited);
    for (int i = 0; i < n; i++) {
      if (es_isLeaf(&P, i)) { iv_push(&leaf, i); continue; }
      if (es_isRoof(&P, i)) {
        int s = P.spouses[i].v[0];
        if (visited[i] > 0 || visited[s] > 0) continue;
        if (sex[i] == 1) { roof.a[roof.n] = i; roof.b[roof.n] = s; }
        else { roof.b[roof.n] = i; roof.a[roof.n] = s; }
        roof.n++;
        visited[i]++; visited[s]++;
        continue;
      }
      if (es_isPeripheral(&P, i)) { iv_push(&peripheral, i); continue; }
    }
  }
  /* BuildPeelingOrder, ES:135-277 */
  int peeled = 0, done = 0;
  for (;;) {
    if (leaf.n == 0 && roof.n == 0 && peripheral.n == 0) break;
    if (done) break;
    while (leaf.n > 0) {
      int aLeaf = leaf.v[0]; iv_del(&leaf, 0);
      peeled++;
      if (P.parents[aLeaf].n < 2) { set_err("peeling error: leaf %d has no parents", aLeaf); rc = PM_EINVAL; goto out; }
      int t0 = P.parents[aLeaf].v[0], t1 = P.parents[aLeaf].v[1];
      steps[nsteps].type = 1; steps[nsteps].from0 = aLeaf; steps[nsteps].from1 = -1;
      steps[nsteps].to0 = t0; steps[nsteps].to1 = t1; nsteps++;
      int idx = iv_find(&P.offspring[t0], aLeaf);
      if (idx < 0) { set_err("Peeling error for person %d", aLeaf); rc = PM_EINVAL; goto out; }
      iv_del(&P.offspring[t0], idx);
      idx = iv_find(&P.offspring[t1], aLeaf);
      if (idx < 0) { set_err("Peeling leaf error: %d", aLeaf); rc = PM_EINVAL; goto out; }
      iv_del(&P.offspring[t1], idx);
      iv_del(&P.parents[aLeaf], 0); iv_del(&P.parents[aLeaf], 0);
      if (es_isPeripheral(&P, t0)) iv_push(&peripheral, t0);
      if (es_isPeripheral(&P, t1)) iv_push(&peripheral, t1);
      int pos = pv_find(&roof, t0, t1);
      if (pos > 0) pv_erase(&roof, pos); /* position 0 is never removed here (ES:185-187) */
      if (peeled == n - 1) done = 1;
    }
    if (done) break;
    while (peripheral.n > 0) {
      int aP = peripheral.v[0]; iv_del(&peripheral, 0);
      peeled++;
      int sp = P.spouses[aP].v[0];
      steps[nsteps].type = 2; steps[nsteps].from0 = aP; steps[nsteps].from1 = -1;
      steps[nsteps].to0 = sp; steps[nsteps].to1 = -1; nsteps++;
      if (P.spouses[aP].n > 1) { set_err("Peripheral parent can not have more than one spouses!"); rc = PM_EINVAL; goto out; }
      int idx = iv_find(&P.spouses[sp], aP);
      if (idx < 0) { set_err("No spouse can be found for person %d", aP); rc = PM_EINVAL; goto out; }
      iv_del(&P.spouses[sp], idx);
      iv_del(&P.spouses[aP], 0);
      if (es_isFinal(&P, sp)) {
        if (peeled != n - 1) { set_err("Are there disconnected sub-pedigrees in the family?"); rc = PM_EINVAL; goto out; }
        done = 1;
        break;
      }
      /* the reference re-reads spouses[aP][0] after deleting it; the stale slot still holds sp */
      if (es_isLeaf(&P, sp)) iv_push(&leaf, sp);
      else if (es_isPeripheral(&P, sp)) iv_push(&peripheral, sp);
      else if (es_isRoof(&P, sp)) es_UpdateRoof(&P, &roof, sp);
    }
    if (done) break;
    if (leaf.n > 0) continue;
    if (peripheral.n > 0) continue;
    while (roof.n > 0) {
      int r0 = roof.a[0], r1 = roof.b[0];
      pv_erase(&roof, 0);
      if (P.offspring[r0].n != 1 || P.offspring[r1].n != 1) { set_err("Roof can only have one offspring for peeling!"); rc = PM_EINVAL; goto out; }
      peeled += 2;
      int child = P.offspring[r0].v[0];
      steps[nsteps].type = 3; steps[nsteps].from0 = r0; steps[nsteps].from1 = r1;
      steps[nsteps].to0 = child; steps[nsteps].to1 = -1; nsteps++;
      iv_del(&P.parents[child], 0); iv_del(&P.parents[child], 0);
      iv_del(&P.offspring[r0], 0); iv_del(&P.offspring[r1], 0);
      if (es_isPeripheral(&P, child)) iv_push(&peripheral, child);
      else if (es_isRoof(&P, child)) es_UpdateRoof(&P, &roof, child);
      else if (es_isFinal(&P, child)) { done = 1; break; }
    }
    if (done) break;
  }
  if (peeled < n - 1) { set_err("Are there inbreeding loops in the pedigree? It cannot handel inbreeding yet!"); rc = PM_EINVAL; }
out:
  free(P.parents); free(P.offspring); free(P.spouses);
  return rc ? rc : nsteps;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct { int key0, key1; double m[10][10]; } mp_t; /* one marriage_partials entry */

typedef struct {
  /* FamilyLikelihoodES state (ES.h:58-96) */
  int famSize, nFounders, first; /* first = index of member 0 in the person arrays */
  int nsteps;
  pm_peel_step *steps;
  int genoIdx[3];
  double (*priors)[10];   /* [nFounders][10] */
  double (*partials)[10]; /* [famSize][10] */
  mp_t mp[32]; int nmp;
} esfam_t;

typedef struct {
  /* NucFamGenotypeLikelihood / ScalarMinimizer state of one famlk[i] object */
  int allele1, allele2, geno11, geno12, geno22;
  double parentPrior[9];
  double (*parentMarginal)[9], (*parentConditional)[9], (*parentGLF)[9]; /* [nFam][9] */
  double a, b, c, min, fa, fb, fc, fmin;
  int isMono, denovo_mono;
  double denovoLR, AB;
  int sex; /* the member `sex` (NucFam.h:72): whatever the last CalcPostProb loop left behind; read by likelihoodONEKid */
  double varPostProb, polyQual;
  double varllk[8], varllk_noprior[8], varfreq[8];
  int totalDepth, numSampWithData;
  double avgDepth, percSampWithData, avgMapQual;
  double (*postProb)[10]; /* [nPerson][10] */
  int *bestGenoIdx; double *dosage; uint8_t *tenState;
  esfam_t *fam; /* [nFam] */
  long n_eval;
} famlk_t;

struct pmo_ctx {
  int nFam, nPerson, nFounders;
  int *famFirst, *famSize, *famFounders, *famGen;
  uint8_t *sex; int *father, *mother;
  pm_params par; /* par.denovo is toggled like main.cpp:569-572 */
  double lut[256];
  double genoMut[10][10];
  double transmission[10][10][10], transmission_denovo[10][10][10], transmission_BA[3][3][3];
  double prior; /* polyPrior, autosome */
  double prior_class[4]; /* GetPolyPrior per PM_CHR_* class (NucFam:231-304) */
  int chrX, chrY, chrMT;  /* SetNonAutosomeFlags of the current site's chromosome (main:312-315) */
  int unrelated;          /* --quick_call pre-pass: MakeUnrelated() is in force (FLSeq:55-60), founders == count everywhere */
  /* current site */
  pm_site_hdr hdr;
  const pm_person_site *ps;
  double (*lk)[10];  /* LUT'ed likelihoods */
  double (*pen)[10]; /* penetrances handed to the ES code (FillPenetrance / FillZeroPenetrance) */
  famlk_t famlk[7];
};

static int fam_isNuclear(const pmo_ctx *c, int f) { return c->famGen[f] == 2 && c->famFounders[f] == 2; }
static int depth_of(const pm_person_site *p) { return p->depth[0] | (p->depth[1] << 8) | (p->depth[2] << 16); }

/* NucFam:89-97 */
static void SetAlleles(famlk_t *k, int a1, int a2) {
  k->allele1 = a1; k->allele2 = a2;
  k->geno11 = GenotypeIndex(a1, a1); k->geno12 = GenotypeIndex(a1, a2); k->geno22 = GenotypeIndex(a2, a2);
}

/* NucFam:383-394 */
static void SetParentPriorSingleTrio(famlk_t *k) {
  static const double t[9] = {0.0, 0.24, 0.04, 0.24, 0.16, 0.08, 0.04, 0.08, 0.12};
  memcpy(k->parentPrior, t, sizeof t);
}
static void hw9(double *pp, double freq) { /* NucFam:323-331 == 372-380 == 410-418 */
  pp[0] = pow(freq, 4);
  pp[1] = freq * freq * freq * (1 - freq) * 2;
  pp[2] = freq * freq * (1 - freq) * (1 - freq);
  pp[3] = freq * (1 - freq) * 2 * freq * freq;
  pp[4] = freq * (1 - freq) * 2 * freq * (1 - freq) * 2;
  pp[5] = freq * (1 - freq) * 2 * (1 - freq) * (1 - freq);
  pp[6] = (1 - freq) * (1 - freq) * freq * freq;
  pp[7] = (1 - freq) * (1 - freq) * freq * (1 - freq) * 2;
  pp[8] = (1 - freq) * (1 - freq) * (1 - freq) * (1 - freq);
}
/* NucFam:318-368 */
static void SetParentPrior(const pmo_ctx *c, famlk_t *k, double freq) {
  if (c->nFam > 1 || k->isMono) {
    double *pp = k->parentPrior;
    if (!c->chrX && !c->chrY && !c->chrMT) hw9(pp, freq);
    if (c->chrX) {
      pp[0] = pow(freq, 3); pp[1] = freq * freq * (1 - freq) * 2; pp[2] = freq * (1 - freq) * (1 - freq);
      pp[3] = 0; pp[4] = 0; pp[5] = 0;
      pp[6] = (1 - freq) * freq * freq; pp[7] = (1 - freq) * freq * (1 - freq) * 2; pp[8] = (1 - freq) * (1 - freq) * (1 - freq);
    }
    if (c->chrY) {
      pp[0] = freq; pp[1] = freq; pp[2] = freq; pp[3] = 0; pp[4] = 0; pp[5] = 0;
      pp[6] = (1 - freq); pp[7] = (1 - freq); pp[8] = (1 - freq);
    }
    if (c->chrMT) {
      pp[0] = freq * freq; pp[1] = 0.0; pp[2] = freq * (1 - freq); pp[3] = 0; pp[4] = 0; pp[5] = 0;
      pp[6] = (1 - freq) * freq; pp[7] = 0; pp[8] = (1 - freq) * (1 - freq);
    }
  } else SetParentPriorSingleTrio(k);
}
/* NucFam:396-420 */
static void SetParentPriorSingleTrio_denovo(famlk_t *k, double freq) {
  if (freq != 1.0) SetParentPriorSingleTrio(k);
  else hw9(k->parentPrior, freq);
}

/* NucFam:1633-1648 */
static void getGenoLikelihood(const pmo_ctx *c, const famlk_t *k, int person, double *lk11, double *lk12, double *lk22) {
  const double *p = c->lk[person];
  *lk11 = p[k->geno11]; *lk12 = p[k->geno12]; *lk22 = p[k->geno22];
}

/* NucFam:1202-1264 ; cfg = 3*gF + gM.  On X/Y/MT the kid's sex is NOT the kid's: the function reads the member
 * `sex` (k->sex), i.e. the last person a CalcPostProb loop of this object visited (0 for famlk[1..6]). */
static double likelihoodONEKid(const pmo_ctx *c, const famlk_t *k, int person, int cfg) {
  double lk = 1.0, lk11, lk12, lk22;
  getGenoLikelihood(c, k, person, &lk11, &lk12, &lk22);
  if (c->chrX || c->chrY || c->chrMT) {
    const int male = k->sex == 1, female = k->sex == 2;
    switch (cfg) {
      case 0: return (c->chrY && female) ? 1.0 : lk11;
      case 1: return c->chrX ? (male ? 0.5 * (lk11 + lk22) : 0.5 * (lk11 + lk12)) : c->chrY ? (male ? lk11 : 1.0) : 0.5 * (lk11 + lk22);
      case 2: return c->chrX ? (male ? lk22 : lk12) : c->chrY ? (male ? lk11 : 1.0) : lk22;
      case 3: case 4: case 5: return 0.0;
      case 6: return c->chrX ? (male ? lk11 : lk12) : c->chrY ? (male ? lk22 : 1.0) : lk11;
      case 7: return c->chrX ? (male ? 0.5 * (lk11 + lk22) : 0.5 * (lk12 + lk22)) : c->chrY ? (male ? lk22 : 1.0) : 0.5 * (lk11 + lk22);
      default: return (c->chrY && female) ? 1.0 : lk22;
    }
  }
  switch (cfg) {
    case 0: lk = lk11; break;
    case 1: lk = 0.5 * (lk11 + lk12); break;
    case 2: lk = lk12; break;
    case 3: lk = 0.5 * (lk11 + lk12); break;
    case 4: lk = 0.25 * lk11 + 0.5 * lk12 + 0.25 * lk22; break;
    case 5: lk = 0.5 * (lk12 + lk22); break;
    case 6: lk = lk12; break;
    case 7: lk = 0.5 * (lk12 + lk22); break;
    case 8: lk = lk22; break;
  }
  return lk;
}
/* NucFam:1184-1198 */
static double likelihoodKids(const pmo_ctx *c, const famlk_t *k, int cfg, int famIdx) {
  double lkKids = 1.0;
  int first = c->famFirst[famIdx], famSize = c->famSize[famIdx];
  for (int i = 2; i < famSize; i++) lkKids *= likelihoodONEKid(c, k, first + i, cfg);
  return lkKids;
}
/* NucFam:1553-1562 */
static double CalcDenovoMutLk(const pmo_ctx *c, const double *ptrlk, int a1, int a2) {
  double lk = 0.0;
  int idx = GenotypeIndex(a1, a2);
  for (int i = 0; i < 10; i++) lk += c->genoMut[idx][i] * ptrlk[i];
  return lk;
}
/* NucFam:1266-1296 */
static double likelihoodONEKid_denovo(const pmo_ctx *c, const famlk_t *k, int person, int cfg) {
  const double *p = c->lk[person];
  int a1 = k->allele1, a2 = k->allele2;
  switch (cfg) {
    case 0: return CalcDenovoMutLk(c, p, a1, a1);
    case 1: return 0.5 * (CalcDenovoMutLk(c, p, a1, a1) + CalcDenovoMutLk(c, p, a1, a2));
    case 2: return CalcDenovoMutLk(c, p, a1, a2);
    case 3: return 0.5 * (CalcDenovoMutLk(c, p, a1, a1) + CalcDenovoMutLk(c, p, a1, a2));
    case 4: return 0.25 * CalcDenovoMutLk(c, p, a1, a1) + 0.5 * CalcDenovoMutLk(c, p, a1, a2) + 0.25 * CalcDenovoMutLk(c, p, a2, a2);
    case 5: return 0.5 * (CalcDenovoMutLk(c, p, a1, a2) + CalcDenovoMutLk(c, p, a2, a2));
    case 6: return CalcDenovoMutLk(c, p, a1, a2);
    case 7: return 0.5 * (CalcDenovoMutLk(c, p, a1, a2) + CalcDenovoMutLk(c, p, a2, a2));
    case 8: return CalcDenovoMutLk(c, p, a2, a2);
  }
  return 1.0;
}
/* NucFam:1299-1312 */
static double likelihoodKids_denovo(const pmo_ctx *c, const famlk_t *k, int cfg, int famIdx) {
  double lkKids = 1.0;
  int first = c->famFirst[famIdx], famSize = c->famSize[famIdx];
  for (int i = 2; i < famSize; i++) lkKids *= likelihoodONEKid_denovo(c, k, first + i, cfg);
  return lkKids;
}

/* NucFam:1046-1061; non_auto: the X/Y/MT lines 1049-1051, which only CalcParentMarginal has (not _denovo) */
static void fill_parentGLF(const pmo_ctx *c, famlk_t *k, int i, int non_auto) {
  double lkF11, lkF12, lkF22, lkM11, lkM12, lkM22;
  int first = c->famFirst[i];
  getGenoLikelihood(c, k, first, &lkF11, &lkF12, &lkF22);
  getGenoLikelihood(c, k, first + 1, &lkM11, &lkM12, &lkM22);
  if (non_auto) {
    if (c->chrX) lkF12 = 0.0;
    if (c->chrY) { lkM11 = lkM12 = lkM22 = 1.0; lkF12 = 0.0; }
    if (c->chrMT) lkF12 = lkM12 = 0.0;
  }
  double *g = k->parentGLF[i];
  g[0] = lkF11 * lkM11; g[1] = lkF11 * lkM12; g[2] = lkF11 * lkM22;
  g[3] = lkF12 * lkM11; g[4] = lkF12 * lkM12; g[5] = lkF12 * lkM22;
  g[6] = lkF22 * lkM11; g[7] = lkF22 * lkM12; g[8] = lkF22 * lkM22;
}
/* NucFam:1041-1084 */
static void CalcParentMarginal(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  fill_parentGLF(c, k, i, 1);
  if (c->nFam > 1 || k->isMono) SetParentPrior(c, k, freq);
  else SetParentPriorSingleTrio(k);
  for (int j = 0; j < 9; j++) k->parentConditional[i][j] = likelihoodKids(c, k, j, i) * k->parentGLF[i][j];
  for (int j = 0; j < 9; j++) k->parentMarginal[i][j] = k->parentConditional[i][j] * k->parentPrior[j];
}
/* NucFam:1086-1132 */
static void CalcParentMarginal_denovo(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  fill_parentGLF(c, k, i, 0);
  if (c->nFam > 1) hw9(k->parentPrior, freq); /* SetParentPrior_denovo */
  else SetParentPriorSingleTrio_denovo(k, freq);
  for (int j = 0; j < 9; j++) k->parentConditional[i][j] = likelihoodKids_denovo(c, k, j, i) * k->parentGLF[i][j];
  for (int j = 0; j < 9; j++) k->parentMarginal[i][j] = k->parentConditional[i][j] * k->parentPrior[j];
}
/* NucFam:987-1004 (autosome) */
static double lkSinglePerson(const pmo_ctx *c, const famlk_t *k, int person, double freq) {
  double sum = 0.0, lk11, lk12, lk22;
  getGenoLikelihood(c, k, person, &lk11, &lk12, &lk22);
  double priors[3];
  priors[0] = freq * freq;
  priors[1] = freq * (1 - freq) * 2;
  priors[2] = (1 - freq) * (1 - freq);
  const int male = c->sex[person] == 1;
  if (c->chrX) { if (male) { lk12 = 0; priors[0] = freq; priors[1] = 0; priors[2] = 1 - freq; } }
  if (c->chrY) { if (male) { lk12 = 0; priors[0] = freq; priors[1] = 0; priors[2] = 1 - freq; } else return 1.0; }
  if (c->chrMT) { lk12 = 0; priors[0] = freq; priors[1] = 0; priors[2] = 1 - freq; }
  sum = sum + lk11 * priors[0] + lk12 * priors[1] + lk22 * priors[2];
  return sum;
}
/* NucFam:941-975 */
static double lkSingleFam(const pmo_ctx *c, famlk_t *k, int i, double freq, int denovo) {
  if (c->unrelated || c->famSize[i] == c->famFounders[i]) {
    const int founders = c->unrelated ? c->famSize[i] : c->famFounders[i];
    double lk = 1.0;
    for (int j = 0; j < founders; j++) lk *= lkSinglePerson(c, k, c->famFirst[i] + j, freq);
    return lk;
  }
  double sum = 0.0;
  if (denovo) CalcParentMarginal_denovo(c, k, i, freq);
  else CalcParentMarginal(c, k, i, freq);
  for (int idx = 0; idx < 9; idx++) sum += k->parentMarginal[i][idx];
  return sum;
}

/* ---- Elston–Stewart ------------------------------------------------------------------------ */
static void es_SetAlleles(esfam_t *e, int a1, int a2) { /* ES:629-635 */
  e->genoIdx[0] = GenotypeIndex(a1, a1); e->genoIdx[1] = GenotypeIndex(a1, a2); e->genoIdx[2] = GenotypeIndex(a2, a2);
}
/* the X/Y/MT founder priors of ES:655-662 == 678-685 on the three slots g0, g1, g2 */
static void es_founder_prior3(const pmo_ctx *c, int male, double freq, double *p0, double *p1, double *p2) {
  *p0 = freq * freq; *p1 = 2 * freq * (1 - freq); *p2 = (1 - freq) * (1 - freq);
  if (c->chrX) if (male) { *p0 = freq; *p1 = 0; *p2 = 1 - freq; }
  if (c->chrY) { if (male) { *p0 = freq; *p1 = 0; *p2 = 1 - freq; } else { *p0 = 1; *p1 = 1; *p2 = 1; } }
  if (c->chrMT) { *p0 = freq; *p1 = 0; *p2 = 1 - freq; }
}
static void es_SetFounderPriors(const pmo_ctx *c, esfam_t *e, double freq) { /* ES:643-664 */
  for (int i = 0; i < e->nFounders; i++) {
    double p0, p1, p2;
    es_founder_prior3(c, c->sex[e->first + i] == 1, freq, &p0, &p1, &p2);
    for (int j = 0; j < 10; j++) e->priors[i][j] = 0.0;
    e->priors[i][e->genoIdx[0]] = p0;
    e->priors[i][e->genoIdx[1]] = p1;
    e->priors[i][e->genoIdx[2]] = p2;
  }
}
static void es_SetFounderPriors_BA(const pmo_ctx *c, esfam_t *e, double freq) { /* ES:666-687 */
  for (int i = 0; i < e->nFounders; i++) {
    for (int j = 0; j < 10; j++) e->priors[i][j] = 0.0;
    es_founder_prior3(c, c->sex[e->first + i] == 1, freq, &e->priors[i][0], &e->priors[i][1], &e->priors[i][2]);
  }
}
static void es_InitializePartials(const pmo_ctx *c, esfam_t *e) { /* ES:1434-1446 */
  for (int i = 0; i < e->famSize; i++) {
    const double *pen = c->pen[e->first + i];
    if (i < e->nFounders) for (int j = 0; j < 10; j++) e->partials[i][j] = e->priors[i][j] * pen[j];
    else for (int j = 0; j < 10; j++) e->partials[i][j] = pen[j];
  }
}
static void es_InitializePartials_BA(const pmo_ctx *c, esfam_t *e) { /* ES:1449-1465 */
  for (int i = 0; i < e->famSize; i++) {
    const double *pen = c->pen[e->first + i];
    const int yfemale = c->chrY && c->sex[e->first + i] == 2;
    for (int j = 0; j < 10; j++) e->partials[i][j] = 0.0;
    if (i < e->nFounders) for (int j = 0; j < 3; j++) e->partials[i][j] = yfemale ? 1.0 : e->priors[i][j] * pen[e->genoIdx[j]];
    else for (int j = 0; j < 3; j++) e->partials[i][j] = yfemale ? 1.0 : pen[e->genoIdx[j]];
  }
}
static mp_t *mp_find(esfam_t *e, int k0, int k1) {
  for (int i = 0; i < e->nmp; i++) if (e->mp[i].key0 == k0 && e->mp[i].key1 == k1) return &e->mp[i];
  return NULL;
}
static mp_t *mp_create(esfam_t *e, int k0, int k1) { /* SetMarriagePartials, ES:1399-1414 */
  mp_t *m = &e->mp[e->nmp++];
  m->key0 = k0; m->key1 = k1;
  for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) m->m[i][j] = 1.0;
  return m;
}
/* GetTransmissionProb_BA, ES:1059-1075, tables ES:834-924 */
static double transmission_BA_of(const pmo_ctx *c, const esfam_t *e, int i, int j, int k, int idx) {
  static const double x2f[27] = {1, 0, 0, .5, .5, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, .5, .5, 0, 0, 1};
  static const double x2m[27] = {1, 0, 0, .5, 0, .5, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, .5, 0, .5, 0, 0, 1};
  static const double chy[27] = {1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1};
  static const double mito[27] = {1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1};
  double transmit = c->transmission_BA[i][j][k];
  const int male = c->sex[e->first + idx] == 1;
  if (c->chrX) transmit = male ? x2m[9 * i + 3 * j + k] : x2f[9 * i + 3 * j + k];
  if (c->chrY) transmit = male ? chy[9 * i + 3 * j + k] : 1.0;
  if (c->chrMT) transmit = mito[9 * i + 3 * j + k];
  return transmit;
}
/* A = 3 uses transmission_BA, A = 10 uses T (= transmission or transmission_denovo) */
static void peelOffspring2Parents(const pmo_ctx *c, esfam_t *e, const pm_peel_step *s, int A, int denovo) { /* ES:1078-1127, 1287-1311 */
  int offspring = s->from0;
  mp_t *m = mp_find(e, s->to0, s->to1);
  if (!m) m = mp_create(e, s->to0, s->to1);
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++) {
      double partial_lk_sum = 0;
      for (int k = 0; k < A; k++) {
        double t = A == 3 ? transmission_BA_of(c, e, i, j, k, offspring) : (denovo ? c->transmission_denovo[i][j][k] : c->transmission[i][j][k]);
        partial_lk_sum += t * e->partials[offspring][k];
      }
      m->m[i][j] *= partial_lk_sum;
    }
}
static void peelSpouse2Spouse(const pmo_ctx *c, esfam_t *e, const pm_peel_step *s, int A) { /* ES:1129-1227, 1313-1361 */
  int spouse_from = s->from0, spouse_to = s->to0;
  int k0, k1, fa2mo;
  if (c->sex[e->first + spouse_from] == 2) { k0 = spouse_to; k1 = spouse_from; fa2mo = 0; }
  else { k0 = spouse_from; k1 = spouse_to; fa2mo = 1; }
  mp_t *m = mp_find(e, k0, k1);
  for (int i = 0; i < A; i++) {
    double partial_lk_sum = 0.0;
    if (!m) for (int j = 0; j < A; j++) partial_lk_sum += e->partials[spouse_from][j];
    else if (fa2mo) for (int j = 0; j < A; j++) partial_lk_sum += e->partials[spouse_from][j] * m->m[j][i];
    else for (int j = 0; j < A; j++) partial_lk_sum += e->partials[spouse_from][j] * m->m[i][j];
    e->partials[spouse_to][i] *= partial_lk_sum;
  }
}
static void peelParents2Offspring(const pmo_ctx *c, esfam_t *e, const pm_peel_step *s, int A, int denovo) { /* ES:1229-1285, 1363-1395 */
  int fa = s->from0, mo = s->from1, offspring = s->to0;
  mp_t *m = mp_find(e, s->from0, s->from1);
  for (int k = 0; k < A; k++) {
    double partial_lk_sum = 0.0;
    if (!m) {
      for (int i = 0; i < A; i++)
        for (int j = 0; j < A; j++) {
          double t = A == 3 ? transmission_BA_of(c, e, i, j, k, offspring) : (denovo ? c->transmission_denovo[i][j][k] : c->transmission[i][j][k]);
          partial_lk_sum += e->partials[fa][i] * e->partials[mo][j] * t;
        }
    } else {
      for (int i = 0; i < A; i++)
        for (int j = 0; j < A; j++) {
          /* de novo with a marriage partial uses the mutation-free tensor (ES:1391) */
          double t = A == 3 ? transmission_BA_of(c, e, i, j, k, offspring) : c->transmission[i][j][k];
          partial_lk_sum += e->partials[fa][i] * m->m[i][j] * e->partials[mo][j] * t;
        }
    }
    e->partials[offspring][k] *= partial_lk_sum;
  }
}
/* ES:990-1057 */
static double es_CalculateLikelihood(const pmo_ctx *c, esfam_t *e, int A, int denovo) {
  e->nmp = 0;
  for (int i = 0; i < e->nsteps; i++) switch (e->steps[i].type) {
    case 1: peelOffspring2Parents(c, e, &e->steps[i], A, denovo); break;
    case 2: peelSpouse2Spouse(c, e, &e->steps[i], A); break;
    case 3: peelParents2Offspring(c, e, &e->steps[i], A, denovo); break;
  }
  int final = e->steps[e->nsteps - 1].to0;
  double lk = 0.0;
  for (int i = 0; i < A; i++) lk += e->partials[final][i];
  return lk;
}
/* FLSeq:256-279 */
static double CalcSingleFamLikelihood_BA(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  esfam_t *e = &k->fam[i];
  es_SetAlleles(e, k->allele1, k->allele2);
  es_SetFounderPriors_BA(c, e, freq);
  es_InitializePartials_BA(c, e);
  return es_CalculateLikelihood(c, e, 3, 0);
}
static double CalcSingleFamLikelihood_denovo(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  esfam_t *e = &k->fam[i];
  es_SetAlleles(e, k->allele1, k->allele2);
  es_SetFounderPriors(c, e, freq);
  es_InitializePartials(c, e);
  return es_CalculateLikelihood(c, e, 10, 1);
}

/* FamilyLikelihoodSeq_VCF::CalcAllFamLogLikelihood, src/FamilyLikelihoodSeq_VCF.cpp:92-109 (autosome):
 * founders-only families sum log10 per person (:111-119), nuclear families use the nuclear formula only when
 * there are several families, everything else goes through the bi-allelic peel. */
static double CalcAllFamLogLikelihood_VCF(pmo_ctx *c, famlk_t *k, double freq) {
  double loglk = 0.0;
  k->n_eval++;
  for (int i = 0; i < c->nFam; i++) {
    if (c->famSize[i] == c->famFounders[i]) {
      double llk = 0.0;
      for (int j = 0; j < c->famSize[i]; j++) llk += log10(lkSinglePerson(c, k, c->famFirst[i] + j, freq));
      loglk += llk;
    } else if (fam_isNuclear(c, i) && c->nFam > 1 && !c->chrX && !c->chrY && !c->chrMT) { /* FLSeq_VCF:101 */
      loglk += log10(lkSingleFam(c, k, i, freq, 0));
    } else {
      loglk += log10(CalcSingleFamLikelihood_BA(c, k, i, freq));
    }
  }
  return loglk;
}

/* FLSeq:222-240, sequential family order (--nthreads 1) */
static double CalcAllFamLogLikelihood(pmo_ctx *c, famlk_t *k, double freq) {
  if (c->par.vcf_input) return CalcAllFamLogLikelihood_VCF(c, k, freq);
  double loglk = 0.0;
  k->n_eval++;
  for (int i = 0; i < c->nFam; i++) {
    if (c->unrelated || fam_isNuclear(c, i) || c->famSize[i] == c->famFounders[i])
      loglk += log10(lkSingleFam(c, k, i, freq, c->par.denovo));
    else
      loglk += c->par.denovo ? log10(CalcSingleFamLikelihood_denovo(c, k, i, freq)) : log10(CalcSingleFamLikelihood_BA(c, k, i, freq));
  }
  return loglk;
}
static double f_obj(pmo_ctx *c, famlk_t *k, double freq) { return -CalcAllFamLogLikelihood(c, k, freq); } /* FLSeq:39-42 */

/* Gold:81-177 */
static double Brent(pmo_ctx *c, famlk_t *k, double tol) {
  double temp;
  if (k->a > k->c) {
    temp = k->a; k->a = k->c; k->c = temp;
    temp = k->fa; k->fa = k->fc; k->fc = temp;
  }
  k->min = k->b; k->fmin = k->fb;
  double w = k->b, v = k->b;
  double fw = k->fb, fv = k->fb;
  double delta = 0.0;
  double u, fu, d = 0.0;
  for (int iter = 1; iter <= ITMAX; iter++) {
    double middle = 0.5 * (k->a + k->c);
    double tol1 = tol * fabs(k->min) + ZEPS;
    double tol2 = 2.0 * tol1;
    if (fabs(k->min - middle) <= (tol2 - 0.5 * (k->c - k->a))) return k->fmin;
    if (fabs(delta) > tol1) {
      double r = (k->min - w) * (k->fmin - fv);
      double q = (k->min - v) * (k->fmin - fw);
      double p = (k->min - v) * q - (k->min - w) * r;
      q = 2.0 * (q - r);
      if (q > 0.0) p = -p;
      q = fabs(q);
      temp = delta;
      delta = d;
      if (fabs(p) >= fabs(0.5 * q * temp) || p <= q * (k->a - k->min) || p >= q * (k->c - k->min)) {
        delta = k->min >= middle ? k->a - k->min : k->c - k->min;
        d = CGOLD * delta;
      } else {
        d = p / q;
        u = k->min + d;
        if (u - k->a < tol2 || k->c - u < tol2) d = sign_d(tol1, middle - k->min);
      }
    } else {
      delta = k->min >= middle ? k->a - k->min : k->c - k->min;
      d = CGOLD * delta;
    }
    u = fabs(d) >= tol1 ? k->min + d : k->min + sign_d(tol1, d);
    fu = f_obj(c, k, u);
    if (fu <= k->fmin) {
      if (u >= k->min) k->a = k->min; else k->c = k->min;
      v = w; w = k->min; k->min = u;
      fv = fw; fw = k->fmin; k->fmin = fu;
    } else {
      if (u < k->min) k->a = u; else k->c = u;
      if (fu <= fw || w == k->min) {
        v = w; w = u;
        fv = fw; fw = fu;
      } else if (fu <= fv || v == k->min || v == w) {
        v = u; fv = fu;
      }
    }
  }
  set_err("ScalarMinimizer::Brent got stuck");
  return k->fmin;
}
/* NucFam:432-444 */
static double OptimizeFrequency(pmo_ctx *c, famlk_t *k) {
  k->a = 0.0001; k->fa = f_obj(c, k, k->a);
  k->b = 0.9999; k->fb = f_obj(c, k, k->b);
  k->c = 0.5;    k->fc = f_obj(c, k, k->c);
  Brent(c, k, c->par.precision);
  return k->min;
}
/* FLSeq:91-104 */
static double PolymorphismLogLikelihood(pmo_ctx *c, famlk_t *k, int a1, int a2) {
  SetAlleles(k, a1, a2);
  /* with founders == count no family isNuclear() (generations == 2 needs a third member): always Brent */
  if (c->unrelated || c->nFam > 1 || (c->nFam == 1 && !fam_isNuclear(c, 0))) {
    OptimizeFrequency(c, k);
    return -k->fmin;
  }
  return CalcAllFamLogLikelihood(c, k, 0.5);
}
/* NucFam:502-517 */
static double MonomorphismLogLikelihood(const pmo_ctx *c, int refBase) {
  double lRef = 0.0;
  int homoRefIdx = GenotypeIndex(refBase, refBase);
  for (int i = 0; i < c->nPerson; i++) lRef += -(double)(c->ps[i].lk[homoRefIdx]) / 10;
  return lRef;
}
/* FLSeq:68-72 */
static double MonomorphismLogLikelihood_denovo(pmo_ctx *c, famlk_t *k, int refBase, int alt) {
  SetAlleles(k, refBase, alt);
  return CalcAllFamLogLikelihood(c, k, 1.0);
}
/* NucFam:520-546 */
static void CalcReadStats(const pmo_ctx *c, famlk_t *k) {
  k->totalDepth = 0; k->numSampWithData = 0; k->avgMapQual = 0.0; k->avgDepth = 0.0;
  for (int i = 0; i < c->nPerson; i++) {
    int d = depth_of(&c->ps[i]);
    k->totalDepth += d;
    k->avgMapQual += c->ps[i].map_quality;
    if (d > 0) k->numSampWithData++;
  }
  if (k->numSampWithData == 0) { k->avgDepth = 0.; k->avgMapQual = 0.; k->percSampWithData = 0.; }
  else {
    k->avgDepth = (double)k->totalDepth / (double)k->numSampWithData;
    k->avgMapQual /= (double)k->numSampWithData;
    k->percSampWithData = (double)k->numSampWithData / (double)c->nPerson;
  }
}
/* NucFam:1664-1683 */
static int CalcMaxLogLkAlt(const famlk_t *k, int refBase, int m) {
  int idx = m;
  double max = k->varllk[m];
  for (int i = m; i < 4; i++) if (max < k->varllk[i]) { max = k->varllk[i]; idx = i; }
  switch (idx) {
    case 0: return refBase;
    case 1: return poly_ts(refBase);
    case 2: return poly_tvs1(refBase);
    default: return poly_tvs2(refBase);
  }
}
/* NucFam:1693-1749 */
static int CalcVarPosterior(famlk_t *k, int refBase, int n) {
  int ts = poly_ts(refBase), tvs1 = poly_tvs1(refBase), tvs2 = poly_tvs2(refBase);
  int maxidx = 0;
  double max = k->varllk[0];
  for (int i = 0; i < n; i++) if (max < k->varllk[i]) { max = k->varllk[i]; maxidx = i; }
  double sumRatio = 0.0;
  for (int i = 0; i < n; i++) sumRatio += pow10(k->varllk[i] - k->varllk[maxidx]);
  k->varPostProb = 1 / sumRatio;
  int a1 = refBase, a2 = refBase;
  switch (maxidx) {
    case 0: a1 = refBase; a2 = CalcMaxLogLkAlt(k, refBase, 1); break;
    case 1: a1 = refBase; a2 = ts; break;
    case 2: a1 = refBase; a2 = tvs1; break;
    case 3: a1 = refBase; a2 = tvs2; break;
    case 4: a1 = ts; a2 = tvs1; break;
    case 5: a1 = ts; a2 = tvs2; break;
    case 6: a1 = tvs1; a2 = tvs2; break;
  }
  SetAlleles(k, a1, a2);
  if (k->varPostProb > 0.9999999999) k->polyQual = 100;
  else k->polyQual = -10 * log10(1 - k->varPostProb);
  return maxidx;
}

/* ---- posteriors ---------------------------------------------------------------------------- */
static int GetBestGenoIdx(double p11, double p12, double p22) { /* NucFam:1564-1571 */
  int bestIdx = 0; double best = p11;
  if (p12 > best) { best = p12; bestIdx = 1; }
  if (p22 > best) { best = p22; bestIdx = 2; }
  return bestIdx;
}
/* NucFam:754-795 (autosome) */
static void CalcPostProb_SinglePerson(const pmo_ctx *c, famlk_t *k, int person, double freq) {
  double lk11, lk12, lk22, priors[3];
  priors[0] = freq * freq; priors[1] = freq * (1 - freq) * 2; priors[2] = (1 - freq) * (1 - freq);
  getGenoLikelihood(c, k, person, &lk11, &lk12, &lk22);
  const int male = c->sex[person] == 1, female = c->sex[person] == 2;
  if (c->chrX && male) { priors[0] = freq; priors[1] = 0.; priors[2] = 1 - freq; }
  if (c->chrY) { if (male) { priors[0] = freq; priors[1] = 0.; priors[2] = 1 - freq; } else { priors[0] = priors[1] = priors[2] = 1.0; } }
  if (c->chrMT) { priors[0] = freq; priors[1] = 0; priors[2] = 1 - freq; }
  double mlk11 = lk11 * priors[0], mlk12 = lk12 * priors[1], mlk22 = lk22 * priors[2];
  double sum = mlk11 + mlk12 + mlk22;
  double *pp = k->postProb[person];
  if (sum == 0) pp[0] = pp[1] = pp[2] = 1 / 3; /* integer division: 0 (NucFam:781) */
  else { pp[0] = mlk11 / sum; pp[1] = mlk12 / sum; pp[2] = mlk22 / sum; }
  if (c->chrY && female) pp[0] = pp[1] = pp[2] = 0.0; /* NucFam:788 */
  k->bestGenoIdx[person] = GetBestGenoIdx(mlk11, mlk12, mlk22);
  k->dosage[person] = pp[1] + pp[2] * 2;
  k->tenState[person] = 0;
}
/* kid config table of NucFam:1334-1443: lk and (lkg11, lkg12, lkg22); `sex` is the kid's own here (NucFam:1349).
 * The all-11 branch has no X/Y/MT case, and the all-22 branch's if/if/if-else chain ends in the autosomal values
 * for everything (MT sets the same ones). */
static void kid_cfg(const pmo_ctx *c, int sex, int cfg, double lk11, double lk12, double lk22, double *lk, double *g11, double *g12, double *g22) {
  if ((c->chrX || c->chrY || c->chrMT) && cfg != 0 && cfg != 8) {
    const int male = sex == 1;
    *g11 = *g12 = *g22 = 0.0;
    if (cfg == 3 || cfg == 4 || cfg == 5) { *lk = 0.0; return; }
    if (c->chrX) {
      switch (cfg) {
        case 1: if (male) { *lk = 0.5 * (lk11 + lk22); *g11 = 0.5 * lk11; *g22 = 0.5 * lk22; } else { *lk = 0.5 * (lk11 + lk12); *g11 = 0.5 * lk11; *g12 = 0.5 * lk12; } break;
        case 2: if (male) { *lk = lk22; *g22 = lk22; } else { *lk = lk12; *g12 = lk12; } break;
        case 6: if (male) { *lk = lk11; *g11 = lk11; } else { *lk = lk12; *g12 = lk12; } break;
        default: if (male) { *lk = 0.5 * (lk11 + lk22); *g11 = 0.5 * lk11; *g22 = 0.5 * lk22; } else { *lk = 0.5 * (lk12 + lk22); *g12 = 0.5 * lk12; *g22 = 0.5 * lk22; } break;
      }
    } else if (c->chrY) {
      if (!male) { *lk = 1.0; return; }
      if (cfg == 1 || cfg == 2) { *lk = lk11; *g11 = lk11; } else { *lk = lk22; *g22 = lk22; }
    } else {
      switch (cfg) {
        case 1: case 7: *lk = 0.5 * (lk11 + lk22); *g11 = 0.5 * lk11; *g22 = 0.5 * lk22; break;
        case 2: *lk = lk22; *g22 = lk22; break;
        default: *lk = lk11; *g11 = lk11; break; /* cfg 6 */
      }
    }
    return;
  }
  switch (cfg) {
    case 0: *lk = lk11; *g11 = lk11; *g12 = *g22 = 0; break;
    case 1: case 3: *lk = 0.5 * (lk11 + lk12); *g11 = lk11 * 0.5; *g12 = lk12 * 0.5; *g22 = 0; break;
    case 2: case 6: *lk = lk12; *g11 = 0; *g12 = lk12; *g22 = 0; break;
    case 4: *lk = 0.25 * lk11 + 0.5 * lk12 + 0.25 * lk22; *g11 = lk11 * 0.25; *g12 = lk12 * 0.5; *g22 = lk22 * 0.25; break;
    case 5: case 7: *lk = 0.5 * (lk12 + lk22); *g11 = 0; *g12 = lk12 * 0.5; *g22 = lk22 * 0.5; break;
    default: *lk = lk22; *g11 = 0; *g12 = 0; *g22 = lk22; break;
  }
}
/* NucFam:590-669 */
static void CalcPostProb_SingleNucFam(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  int first = c->famFirst[i], famSize = c->famSize[i];
  if (famSize <= c->famFounders[i]) {
    for (int j = 0; j < c->famFounders[i]; j++) { k->sex = c->sex[first + j]; CalcPostProb_SinglePerson(c, k, first + j, freq); }
    return;
  }
  CalcParentMarginal(c, k, i, freq); /* on X/Y/MT: with the `sex` the previous family's loop left (NucFam:606 vs 610) */
  const double *pm = k->parentMarginal[i];
  for (int j = 0; j < famSize; j++) {
    double p11, p12, p22, sum;
    double *pp = k->postProb[first + j];
    k->tenState[first + j] = 0;
    k->sex = c->sex[first + j];
    if (j == 0 || j == 1) {
      if (j == 0) { p11 = pm[0] + pm[1] + pm[2]; p12 = pm[3] + pm[4] + pm[5]; p22 = pm[6] + pm[7] + pm[8]; }
      else { p11 = pm[0] + pm[3] + pm[6]; p12 = pm[1] + pm[4] + pm[7]; p22 = pm[2] + pm[5] + pm[8]; }
      sum = p11 + p12 + p22;
      if (sum == 0) pp[0] = pp[1] = pp[2] = 1 / 3;
      else { pp[0] = p11 / sum; pp[1] = p12 / sum; pp[2] = p22 / sum; }
      k->bestGenoIdx[first + j] = GetBestGenoIdx(p11, p12, p22);
    } else {
      /* KidJointGenoLikelihood NucFam:798-835 + likelihoodKidGenotype 1334-1443 */
      double J11 = 0, J12 = 0, J22 = 0;
      for (int cfg = 0; cfg < 9; cfg++) {
        double G11 = 1.0, G12 = 1.0, G22 = 1.0;
        for (int kk = 2; kk < famSize; kk++) {
          double lk11, lk12, lk22, lk, g11, g12, g22;
          getGenoLikelihood(c, k, first + kk, &lk11, &lk12, &lk22);
          kid_cfg(c, c->sex[first + kk], cfg, lk11, lk12, lk22, &lk, &g11, &g12, &g22);
          if (kk != j) { G11 *= lk; G12 *= lk; G22 *= lk; }
          else { G11 *= g11; G12 *= g12; G22 *= g22; }
        }
        double w = k->parentGLF[i][cfg] * k->parentPrior[cfg];
        G11 *= w; G12 *= w; G22 *= w;
        if (cfg == 0) { J11 = G11; J12 = G12; J22 = G22; }
        else { J11 = J11 + G11; J12 = J12 + G12; J22 = J22 + G22; }
      }
      sum = J11 + J12 + J22;
      double post11 = 0, post12 = 0, post22 = 0; /* JointGenoLk::CalcPost PedigreeGLF.cpp:12-21 */
      if (sum != 0.0) { post11 = J11 / sum; post12 = J12 / sum; post22 = J22 / sum; }
      pp[0] = post11; pp[1] = post12; pp[2] = post22;
      k->bestGenoIdx[first + j] = GetBestGenoIdx(post11, post12, post22);
    }
    k->dosage[first + j] = pp[1] + pp[2] * 2;
  }
}
/* GetJointGenoLk_denovo NucFam:1480-1551 */
static void GetJointGenoLk_denovo(const pmo_ctx *c, const famlk_t *k, int person, int cfg, double *out) {
  const double *p = c->lk[person];
  const double *M1 = c->genoMut[GenotypeIndex(k->allele1, k->allele1)];
  const double *M2 = c->genoMut[GenotypeIndex(k->allele1, k->allele2)];
  const double *M3 = c->genoMut[GenotypeIndex(k->allele2, k->allele2)];
  for (int i = 0; i < 10; i++) switch (cfg) {
    case 0: out[i] = M1[i] * p[i]; break;
    case 1: case 3: out[i] = (0.5 * M1[i] + 0.5 * M2[i]) * p[i]; break;
    case 2: case 6: out[i] = M2[i] * p[i]; break;
    case 4: out[i] = (0.25 * M1[i] + 0.5 * M2[i] + 0.25 * M3[i]) * p[i]; break;
    case 5: case 7: out[i] = (0.5 * M2[i] + 0.5 * M3[i]) * p[i]; break;
    default: out[i] = M3[i] * p[i]; break;
  }
}
/* NucFam:671-752 */
static void CalcPostProb_SingleNucFam_denovo(const pmo_ctx *c, famlk_t *k, int i, double freq) {
  int first = c->famFirst[i], famSize = c->famSize[i];
  if (famSize <= c->famFounders[i]) {
    for (int j = 0; j < c->famFounders[i]; j++) CalcPostProb_SinglePerson(c, k, first + j, freq);
    return;
  }
  CalcParentMarginal_denovo(c, k, i, freq);
  const double *pm = k->parentMarginal[i];
  for (int j = 0; j < famSize; j++) {
    double *pp = k->postProb[first + j];
    if (j < 2) {
      double p11, p12, p22, sum;
      if (j == 0) { p11 = pm[0] + pm[1] + pm[2]; p12 = pm[3] + pm[4] + pm[5]; p22 = pm[6] + pm[7] + pm[8]; }
      else { p11 = pm[0] + pm[3] + pm[6]; p12 = pm[1] + pm[4] + pm[7]; p22 = pm[2] + pm[5] + pm[8]; }
      sum = p11 + p12 + p22;
      if (sum == 0) pp[0] = pp[1] = pp[2] = 1 / 3;
      else { pp[0] = p11 / sum; pp[1] = p12 / sum; pp[2] = p22 / sum; }
      k->bestGenoIdx[first + j] = GetBestGenoIdx(p11, p12, p22);
      k->dosage[first + j] = pp[1] + pp[2] * 2;
      k->tenState[first + j] = 0;
    } else {
      /* KidJointGenoLikelihood_denovo NucFam:838-868, likelihoodKidGenotype_denovo 1446-1478 */
      double geno[10];
      for (int g = 0; g < 10; g++) geno[g] = 0.0;
      for (int cfg = 0; cfg < 9; cfg++) {
        double lkKidGeno[10];
        for (int g = 0; g < 10; g++) lkKidGeno[g] = 1.0;
        for (int kk = 2; kk < famSize; kk++) {
          if (kk != j) {
            double lk = likelihoodONEKid_denovo(c, k, first + kk, cfg);
            for (int g = 0; g < 10; g++) lkKidGeno[g] *= lk;
          } else {
            double joint[10];
            GetJointGenoLk_denovo(c, k, first + kk, cfg, joint);
            for (int g = 0; g < 10; g++) lkKidGeno[g] *= joint[g];
          }
        }
        double w = k->parentGLF[i][cfg] * k->parentPrior[cfg];
        for (int g = 0; g < 10; g++) lkKidGeno[g] *= w;
        for (int g = 0; g < 10; g++) geno[g] += lkKidGeno[g];
      }
      double sum = 0.0;
      for (int g = 0; g < 10; g++) sum += geno[g];
      if (sum == 0.0) for (int g = 0; g < 10; g++) pp[g] = 0.0;
      else for (int g = 0; g < 10; g++) pp[g] = geno[g] / sum;
      double maxPost = 0.0; int best = 0;
      for (int g = 0; g < 10; g++) if (maxPost < pp[g]) { maxPost = pp[g]; best = g; }
      k->bestGenoIdx[first + j] = best;
      k->dosage[first + j] = 0.0;
      k->tenState[first + j] = 1;
    }
  }
}
/* FLSeq:327-356: penetrances with person `person` pinned to genotype genoIdx */
static void FillZeroPenetrance(pmo_ctx *c, int famIdx, int person, int genoIdx) {
  int first = c->famFirst[famIdx];
  for (int i = 0; i < c->famSize[famIdx]; i++)
    for (int j = 0; j < 10; j++)
      c->pen[first + i][j] = (i != person || j == genoIdx) ? c->lk[first + i][j] : 0.0;
}
static void FillPenetrance(pmo_ctx *c) { /* FLSeq:296-317 */
  memcpy(c->pen, c->lk, sizeof(double) * 10 * (size_t)c->nPerson);
}
/* FLSeq:182-216 */
static void CalcPostProb_SingleExtendedPed_BA(pmo_ctx *c, famlk_t *k, int i, double freq) {
  int first = c->famFirst[i];
  for (int j = 0; j < c->famSize[i]; j++) {
    double lk11, lk12, lk22, sum;
    k->sex = c->sex[first + j];
    if (c->chrY && c->sex[first + j] == 2) { /* FLSeq:181-188 */
      double *pp = k->postProb[first + j];
      k->bestGenoIdx[first + j] = 0; pp[0] = pp[1] = pp[2] = 0.0; k->dosage[first + j] = 0; k->tenState[first + j] = 0;
      continue;
    }
    FillZeroPenetrance(c, i, j, GenotypeIndex(k->allele1, k->allele1));
    lk11 = CalcSingleFamLikelihood_BA(c, k, i, freq);
    FillZeroPenetrance(c, i, j, GenotypeIndex(k->allele1, k->allele2));
    lk12 = CalcSingleFamLikelihood_BA(c, k, i, freq);
    FillZeroPenetrance(c, i, j, GenotypeIndex(k->allele2, k->allele2));
    lk22 = CalcSingleFamLikelihood_BA(c, k, i, freq);
    sum = lk11 + lk12 + lk22;
    double *pp = k->postProb[first + j];
    if (sum == 0) pp[0] = pp[1] = pp[2] = 0.0;
    else { pp[0] = lk11 / sum; pp[1] = lk12 / sum; pp[2] = lk22 / sum; }
    k->bestGenoIdx[first + j] = GetBestGenoIdx(lk11, lk12, lk22);
    k->dosage[first + j] = pp[1] + pp[2] * 2;
    k->tenState[first + j] = 0;
  }
  FillPenetrance(c);
}
/* FLSeq:140-180 */
static void CalcPostProb_SingleExtendedPed_denovo(pmo_ctx *c, famlk_t *k, int i, double freq) {
  int first = c->famFirst[i];
  for (int j = 0; j < c->famSize[i]; j++) {
    double lk[10], sum = 0.0;
    for (int g = 0; g < 10; g++) {
      FillZeroPenetrance(c, i, j, g);
      lk[g] = CalcSingleFamLikelihood_denovo(c, k, i, freq);
    }
    for (int g = 0; g < 10; g++) sum += lk[g];
    double *pp = k->postProb[first + j];
    if (sum == 0) for (int g = 0; g < 10; g++) pp[g] = 0;
    else for (int g = 0; g < 10; g++) pp[g] = lk[g] / sum;
    int best = 0; double max = 0.0;
    for (int g = 0; g < 10; g++) if (max < lk[g]) { max = lk[g]; best = g; }
    k->bestGenoIdx[first + j] = best;
    /* dosage[][] is not touched on this path (FLSeq:140-180): it keeps its previous value; the
       de novo writer never prints DS.  We report 0. */
    k->dosage[first + j] = 0.0;
    k->tenState[first + j] = 1;
  }
  FillPenetrance(c);
}
/* FLSeq:74-89 */
static void CalcPostProb(pmo_ctx *c, famlk_t *k, double freq) {
  for (int i = 0; i < c->nFam; i++) {
    if (fam_isNuclear(c, i) || c->famSize[i] == c->famFounders[i]) {
      if (c->par.denovo) CalcPostProb_SingleNucFam_denovo(c, k, i, freq);
      else CalcPostProb_SingleNucFam(c, k, i, freq);
    } else {
      if (c->par.denovo) CalcPostProb_SingleExtendedPed_denovo(c, k, i, freq);
      else CalcPostProb_SingleExtendedPed_BA(c, k, i, freq);
    }
  }
}
/* NucFam:1006-1039 */
static void CalculateAB(const pmo_ctx *c, famlk_t *k, double freq) {
  k->AB = 0.5;
  double A = 0.0, B = 0.0;
  double p11 = freq * freq, p12 = 2 * freq * (1 - freq), p22 = (1 - freq) * (1 - freq);
  for (int i = 0; i < c->nPerson; i++) {
    int depth = depth_of(&c->ps[i]);
    double lk11, lk12, lk22;
    getGenoLikelihood(c, k, i, &lk11, &lk12, &lk22);
    unsigned char llk11 = c->ps[i].lk[k->geno11], llk12 = c->ps[i].lk[k->geno12], llk22 = c->ps[i].lk[k->geno22];
    double PHet = (p12 * lk12) / (p11 * lk11 + p12 * lk12 + p22 * lk22);
    if (PHet > 1e-05 && depth > 0) {
      int scale = llk22 + llk11 - 2 * llk12 + 6 * depth;
      int minimum = abs(llk22 - llk11);
      if (scale < 4) scale = 4;
      if (scale < minimum) scale = minimum;
      int nRef = 0.5 * depth * (1 + (llk22 - llk11) / (scale + 1e-30));
      A += PHet * nRef;
      B += PHet * depth;
    }
  }
  k->AB = (0.05 + A) / (0.1 + B);
}

/* ---- context ------------------------------------------------------------------------------ */
static void build_transmission(pmo_ctx *c) {
  /* ES:752-785 */
  memset(c->transmission, 0, sizeof c->transmission);
  for (int i = 1; i <= 4; i++)
    for (int j = i; j <= 4; j++) {
      int idx1 = GenotypeIndex(i, j);
      for (int k = 1; k <= 4; k++)
        for (int m = k; m <= 4; m++) {
          int idx2 = GenotypeIndex(k, m);
          int geno[4] = {GenotypeIndex(i, k), GenotypeIndex(i, m), GenotypeIndex(j, k), GenotypeIndex(j, m)};
          for (int t = 0; t < 4; t++) c->transmission[idx1][idx2][geno[t]] += 0.25;
        }
    }
  /* ES:787-810 */
  for (int i = 0; i < 10; i++)
    for (int j = 0; j < 10; j++)
      for (int k = 0; k < 10; k++) {
        double sum = .0;
        for (int m = 0; m < 10; m++) sum += c->transmission[i][j][m] * c->genoMut[m][k];
        c->transmission_denovo[i][j][k] = sum;
      }
  /* ES:812-832 */
  static const double tba[27] = {1, 0, 0, .5, .5, 0, 0, 1, 0, .5, .5, 0, .25, .5, .25, 0, .5, .5, 0, 1, 0, 0, .5, .5, 0, 0, 1};
  memcpy(c->transmission_BA, tba, sizeof tba);
}

pmo_ctx *pmo_create(const pm_pedigree *ped, const pm_params *par, const double *lut256) {
  pmo_ctx *c = calloc(1, sizeof *c);
  c->nFam = ped->n_fam; c->nPerson = ped->n_person;
  c->famFirst = calloc((size_t)c->nFam, sizeof(int)); c->famSize = calloc((size_t)c->nFam, sizeof(int));
  c->famFounders = calloc((size_t)c->nFam, sizeof(int)); c->famGen = calloc((size_t)c->nFam, sizeof(int));
  c->sex = calloc((size_t)c->nPerson, 1);
  c->father = calloc((size_t)c->nPerson, sizeof(int)); c->mother = calloc((size_t)c->nPerson, sizeof(int));
  int off = 0;
  for (int f = 0; f < c->nFam; f++) {
    c->famFirst[f] = off; c->famSize[f] = ped->fam_size[f];
    c->famFounders[f] = ped->fam_founders[f]; c->famGen[f] = ped->fam_generations[f];
    off += ped->fam_size[f]; c->nFounders += ped->fam_founders[f];
  }
  if (off != c->nPerson) { set_err("n_person != sum(fam_size)"); pmo_destroy(c); return NULL; }
  memcpy(c->sex, ped->sex, (size_t)c->nPerson);
  memcpy(c->father, ped->father, sizeof(int) * (size_t)c->nPerson);
  memcpy(c->mother, ped->mother, sizeof(int) * (size_t)c->nPerson);
  c->par = *par;
  if (lut256) memcpy(c->lut, lut256, sizeof c->lut); else pmo_fill_lut(c->lut);
  pmo_genotype_mutation_matrix(par->denovo_mut_rate, par->denovo_tstv, &c->genoMut[0][0]);
  build_transmission(c);
  /* SetPolyPrior NucFam:231-242 */
  if (c->nFounders == 0) { set_err("Family size is zero"); pmo_destroy(c); return NULL; }
  c->prior = 0;
  for (int i = 1; i <= 2 * c->nFounders; i++) c->prior += 1.0 / i;
  c->prior *= par->theta;
  { /* SetPolyPrior_chrX / _chrY / _MT, NucFam:256-293: founder chromosomes by sex */
    int maleFounders = 0, femaleFounders = 0;
    for (int i = 0; i < c->nPerson; i++)
      if (c->father[i] < 0 && c->mother[i] < 0) { if (c->sex[i] == 1) maleFounders++; if (c->sex[i] == 2) femaleFounders++; }
    const int n_chr[4] = {2 * c->nFounders, femaleFounders * 2 + maleFounders, maleFounders, c->nFounders};
    for (int cl = 0; cl < 4; cl++) {
      double pr = 0;
      for (int i = 1; i <= n_chr[cl]; i++) pr += 1.0 / i;
      c->prior_class[cl] = pr * par->theta;
    }
  }
  c->lk = calloc((size_t)c->nPerson, sizeof *c->lk);
  c->pen = calloc((size_t)c->nPerson, sizeof *c->pen);
  for (int r = 0; r < 7; r++) {
    famlk_t *k = &c->famlk[r];
    k->parentMarginal = calloc((size_t)c->nFam, sizeof *k->parentMarginal);
    k->parentConditional = calloc((size_t)c->nFam, sizeof *k->parentConditional);
    k->parentGLF = calloc((size_t)c->nFam, sizeof *k->parentGLF);
    k->postProb = calloc((size_t)c->nPerson, sizeof *k->postProb);
    k->bestGenoIdx = calloc((size_t)c->nPerson, sizeof(int));
    k->dosage = calloc((size_t)c->nPerson, sizeof(double));
    k->tenState = calloc((size_t)c->nPerson, 1);
    k->denovoLR = -1; k->AB = 0.5;
    k->fam = calloc((size_t)c->nFam, sizeof(esfam_t));
    for (int f = 0; f < c->nFam; f++) {
      esfam_t *e = &k->fam[f];
      e->famSize = c->famSize[f]; e->nFounders = c->famFounders[f]; e->first = c->famFirst[f];
      e->priors = calloc((size_t)(e->nFounders > 0 ? e->nFounders : 1), sizeof *e->priors);
      e->partials = calloc((size_t)e->famSize, sizeof *e->partials);
      if (e->famSize != e->nFounders && (!fam_isNuclear(c, f) || par->vcf_input)) {
        e->steps = calloc((size_t)e->famSize, sizeof(pm_peel_step));
        int n = pmo_build_peel_order(e->famSize, c->father + e->first, c->mother + e->first, c->sex + e->first, e->steps);
        if (n < 0) { pmo_destroy(c); return NULL; }
        e->nsteps = n;
      }
    }
  }
  return c;
}

void pmo_destroy(pmo_ctx *c) {
  if (!c) return;
  for (int r = 0; r < 7; r++) {
    famlk_t *k = &c->famlk[r];
    if (k->fam) for (int f = 0; f < c->nFam; f++) { free(k->fam[f].priors); free(k->fam[f].partials); free(k->fam[f].steps); }
    free(k->fam); free(k->parentMarginal); free(k->parentConditional); free(k->parentGLF);
    free(k->postProb); free(k->bestGenoIdx); free(k->dosage); free(k->tenState);
  }
  free(c->famFirst); free(c->famSize); free(c->famFounders); free(c->famGen);
  free(c->sex); free(c->father); free(c->mother); free(c->lk); free(c->pen);
  free(c);
}

int pmo_load_site(pmo_ctx *c, const pm_site_hdr *hdr, const pm_person_site *persons) {
  c->hdr = *hdr; c->ps = persons;
  c->chrX = hdr->chr_class == PM_CHR_X; c->chrY = hdr->chr_class == PM_CHR_Y; c->chrMT = hdr->chr_class == PM_CHR_MT;
  for (int i = 0; i < c->nPerson; i++)
    for (int j = 0; j < 10; j++) c->lk[i][j] = c->lut[persons[i].lk[j]]; /* glfHandler.cpp:230-231 */
  FillPenetrance(c);
  return PM_OK;
}
double pmo_family_loglik(pmo_ctx *c, int fam, int a1, int a2, double freq, int denovo) {
  famlk_t *k = &c->famlk[1];
  int save = c->par.denovo; c->par.denovo = denovo;
  SetAlleles(k, a1, a2);
  double r;
  if (fam_isNuclear(c, fam) || c->famSize[fam] == c->famFounders[fam]) r = log10(lkSingleFam(c, k, fam, freq, denovo));
  else r = denovo ? log10(CalcSingleFamLikelihood_denovo(c, k, fam, freq)) : log10(CalcSingleFamLikelihood_BA(c, k, fam, freq));
  c->par.denovo = save;
  return r;
}
double pmo_all_family_loglik(pmo_ctx *c, int a1, int a2, double freq, int denovo) {
  famlk_t *k = &c->famlk[1];
  int save = c->par.denovo; c->par.denovo = denovo;
  SetAlleles(k, a1, a2);
  double r = CalcAllFamLogLikelihood(c, k, freq);
  c->par.denovo = save;
  return r;
}
double pmo_optimize(pmo_ctx *c, int a1, int a2, int denovo, double *freq, int *n_eval) {
  famlk_t *k = &c->famlk[1];
  int save = c->par.denovo; c->par.denovo = denovo;
  SetAlleles(k, a1, a2);
  long e0 = k->n_eval;
  OptimizeFrequency(c, k);
  if (freq) *freq = k->min;
  if (n_eval) *n_eval = (int)(k->n_eval - e0);
  c->par.denovo = save;
  return -k->fmin;
}

/* ---- the per-site control loop, main:325-594 ------------------------------------------------- */
static void fill_result(const pmo_ctx *c, const famlk_t *k, pm_site_result *r) {
  r->allele1 = (uint8_t)k->allele1; r->allele2 = (uint8_t)k->allele2;
  r->total_depth = k->totalDepth; r->num_samp = k->numSampWithData;
  r->perc_samp = k->percSampWithData; r->avg_map_qual = k->avgMapQual;
  r->var_post_prob = k->varPostProb; r->poly_qual = k->polyQual;
  r->freq = k->min; r->denovo_lr = k->denovoLR; r->ab = k->AB;
  /* slots of hypotheses that were not evaluated at this site hold stale values from earlier sites in
     the reference's members; they are never read, so the result reports them as 0 */
  for (int i = 0; i < 7; i++) {
    int live = i < r->n_hyp;
    r->varllk[i] = live ? k->varllk[i] : 0.0;
    r->varllk_noprior[i] = live ? k->varllk_noprior[i] : 0.0;
    r->varfreq[i] = live ? k->varfreq[i] : 0.0;
  }
  (void)c;
}

static int call_site(pmo_ctx *c, uint32_t site, pm_site_result *r, pm_person_result *pr) {
  famlk_t *fl = c->famlk;
  const pm_params *par = &c->par;
  memset(r, 0, sizeof *r);
  r->site = site; r->maxidx = -1;
  int refBase = c->hdr.ref_base;
  if (refBase != 1 && refBase != 2 && refBase != 3 && refBase != 4) { r->status = PM_SITE_BAD_REF; return 0; }
  if (c->hdr.chr_class > PM_CHR_MT) { set_err("oracle: bad chr_class %d", c->hdr.chr_class); return PM_EINVAL; }
  double polyPrior = c->prior_class[c->hdr.chr_class];
  /* The member `sex` of the seven objects (quirk 6 of SURVEY 8a).  famlk[1..6] never run CalcPostProb: 0.  famlk[0]:
     without --denovo every CalcPostProb leaves the sex of the last person of the last family, so that is what the next
     emitted site's first family sees -- except in the very first CalcPostProb of the process (hdr.reserved bit 0, set
     by the caller for that one site).  With --denovo no loop ever assigns it. */
  for (int r7 = 1; r7 < 7; r7++) fl[r7].sex = 0;
  fl[0].sex = (par->denovo || (c->hdr.reserved & PM_HDR_FIRST_POSTPROB)) ? 0 : c->sex[c->nPerson - 1];
  double prior_ts = par->poly_tstv / (par->poly_tstv + 1); /* main:192-193 */
  double prior_tv = (1 - prior_ts) / 2;

  CalcReadStats(c, &fl[0]);
  r->total_depth = fl[0].totalDepth; r->num_samp = fl[0].numSampWithData;
  r->perc_samp = fl[0].percSampWithData; r->avg_map_qual = fl[0].avgMapQual;
  if (fl[0].totalDepth < par->min_total_depth) { r->status = PM_SITE_MIN_DEPTH; return 0; }
  if (par->max_total_depth > 0 && fl[0].totalDepth > par->max_total_depth) { r->status = PM_SITE_MAX_DEPTH; return 0; }
  if (fl[0].percSampWithData * 100 < par->min_ps) { r->status = PM_SITE_MIN_PS; return 0; }
  if (fl[0].avgMapQual < par->min_map_quality) { r->status = PM_SITE_MIN_MAPQ; return 0; }

  int ts = poly_ts(refBase), tvs1 = poly_tvs1(refBase), tvs2 = poly_tvs2(refBase);
  if (par->quick_call) { /* main:354-437: the same four (seven) hypotheses with everybody unrelated */
    c->unrelated = 1;
    const double polyPrior_unr = polyPrior; /* GetPolyPrior_unr() == GetPolyPrior(): both read nFounders (NucFam:295-311) */
    fl[0].varllk[0] = log10(1 - polyPrior_unr) + MonomorphismLogLikelihood(c, refBase);
    fl[0].varllk[1] = log10(polyPrior_unr * prior_ts) + PolymorphismLogLikelihood(c, &fl[1], refBase, ts);
    fl[0].varllk[2] = log10(polyPrior_unr * prior_tv) + PolymorphismLogLikelihood(c, &fl[2], refBase, tvs1);
    fl[0].varllk[3] = log10(polyPrior_unr * prior_tv) + PolymorphismLogLikelihood(c, &fl[3], refBase, tvs2);
    int qidx = CalcVarPosterior(&fl[0], refBase, 4);
    if (fl[0].varPostProb < 0.99) {
      fl[0].varllk[4] = log10(polyPrior_unr * 0.001) + PolymorphismLogLikelihood(c, &fl[4], ts, tvs1);
      fl[0].varllk[5] = log10(polyPrior_unr * 0.001) + PolymorphismLogLikelihood(c, &fl[5], ts, tvs2);
      fl[0].varllk[6] = log10(polyPrior_unr * 0.001) + PolymorphismLogLikelihood(c, &fl[6], tvs1, tvs2);
      qidx = CalcVarPosterior(&fl[0], refBase, 7);
    }
    /* the two `continue`s below skip RestoreFounderCount() in the reference; the next site's pre-pass calls
       MakeUnrelated() again and nothing in between reads the founder counts, so the leak is not observable */
    c->unrelated = 0;
    if (fl[0].varPostProb < par->posterior_cutoff || qidx == 0) { r->status = PM_SITE_QUICK_SKIP; return 0; }
    /* A lone nuclear family does not run Brent in the real pass (FLSeq:94-103), so in the reference famlk[1..6].min
       keep the pre-pass optima (or older sites' ones) and leak into famlk[0].min.  Nothing prints it there (no AF, no
       AB for a single nuclear family, NucFam:1795-1802); the restatement reports the same 0 as without --quick_call. */
    if (c->nFam == 1 && fam_isNuclear(c, 0)) for (int h = 1; h < 7; h++) fl[h].min = 0.0;
  }

  /* main:439-495 */
  if (!par->denovo) {
    double lRef = log10(1 - polyPrior) + MonomorphismLogLikelihood(c, refBase);
    fl[0].varllk[0] = lRef;
    fl[0].varllk_noprior[0] = lRef - log10(1 - polyPrior);
    fl[0].varfreq[0] = 1.0;
  } else {
    double lRef_denovo = log10(1 - polyPrior) + MonomorphismLogLikelihood_denovo(c, &fl[0], refBase, refBase == 4 ? refBase - 1 : refBase + 1);
    fl[0].varllk[0] = lRef_denovo;
    fl[0].varllk_noprior[0] = lRef_denovo - log10(1 - polyPrior);
    fl[0].varfreq[0] = 1.0;
  }
  {
    double v = log10(polyPrior * prior_ts) + PolymorphismLogLikelihood(c, &fl[1], refBase, ts);
    fl[0].varllk[1] = v; fl[0].varllk_noprior[1] = v - log10(polyPrior * 2. / 3.); fl[0].varfreq[1] = fl[1].min;
    v = log10(polyPrior * prior_tv) + PolymorphismLogLikelihood(c, &fl[2], refBase, tvs1);
    fl[0].varllk[2] = v; fl[0].varllk_noprior[2] = v - log10(polyPrior * 1. / 6.); fl[0].varfreq[2] = fl[2].min;
    v = log10(polyPrior * prior_tv) + PolymorphismLogLikelihood(c, &fl[3], refBase, tvs2);
    fl[0].varllk[3] = v; fl[0].varllk_noprior[3] = v - log10(polyPrior * 1. / 6.); fl[0].varfreq[3] = fl[3].min;
  }
  int maxidx = CalcVarPosterior(&fl[0], refBase, 4);
  r->n_hyp = 4;
  if (fl[0].varPostProb < 0.99) { /* main:499-537 */
    double v = log10(polyPrior * 0.001) + PolymorphismLogLikelihood(c, &fl[4], ts, tvs1);
    fl[0].varllk[4] = v; fl[0].varllk_noprior[4] = v - log10(polyPrior * 0.001); fl[0].varfreq[4] = fl[4].min;
    v = log10(polyPrior * 0.001) + PolymorphismLogLikelihood(c, &fl[5], ts, tvs2);
    fl[0].varllk[5] = v; fl[0].varllk_noprior[5] = v - log10(polyPrior * 0.001); fl[0].varfreq[5] = fl[5].min;
    v = log10(polyPrior * 0.001) + PolymorphismLogLikelihood(c, &fl[6], tvs1, tvs2);
    fl[0].varllk[6] = v; fl[0].varllk_noprior[6] = v - log10(polyPrior * 0.001); fl[0].varfreq[6] = fl[6].min;
    maxidx = CalcVarPosterior(&fl[0], refBase, 7);
    r->n_hyp = 7;
  }
  r->maxidx = (int8_t)maxidx;
  fill_result(c, &fl[0], r);

  if (fl[0].varPostProb < par->posterior_cutoff) { /* main:539 */
    r->flags |= PM_FLAG_NOCALL;
    if (!par->force_call && !par->out_all_sites) { r->status = PM_SITE_NOCALL; return 0; }
  }
  switch (maxidx) { /* main:541-553 */
    case 0: if (par->force_call || par->out_all_sites) fl[0].min = 1.0; break;
    case 1: SetAlleles(&fl[0], refBase, ts); fl[0].min = fl[1].min; break;
    case 2: SetAlleles(&fl[0], refBase, tvs1); fl[0].min = fl[2].min; break;
    case 3: SetAlleles(&fl[0], refBase, tvs2); fl[0].min = fl[3].min; break;
    case 4: SetAlleles(&fl[0], ts, tvs1); fl[0].min = fl[4].min; break;
    case 5: SetAlleles(&fl[0], ts, tvs2); fl[0].min = fl[5].min; break;
    case 6: SetAlleles(&fl[0], tvs1, tvs2); fl[0].min = fl[6].min; break;
  }
  if (maxidx == 0 && par->denovo == 0 && par->force_call == 0 && par->out_all_sites == 0) { /* main:555 */
    fill_result(c, &fl[0], r);
    r->status = PM_SITE_MONO; return 0;
  }
  if (maxidx == 0) { /* main:557-565 */
    if (par->denovo) {
      double lk_mono = MonomorphismLogLikelihood(c, refBase);
      fl[0].min = 1.0;
      fl[0].denovoLR = fl[0].varllk_noprior[0] - lk_mono;
      if (fl[0].denovoLR <= log10(par->denovo_min_llr) && !par->out_all_sites && !par->force_call) {
        fill_result(c, &fl[0], r);
        r->status = PM_SITE_DENOVO_LOW_LR; return 0;
      }
    }
  } else if (par->denovo) { /* main:566-574 */
    c->par.denovo = 0;
    double lk_poly = PolymorphismLogLikelihood(c, &fl[0], fl[0].allele1, fl[0].allele2);
    fl[0].denovoLR = fl[0].varllk_noprior[maxidx] - lk_poly;
    c->par.denovo = 1;
    r->refit_llk = lk_poly;
  }
  /* OutputVCF_denovo returns before printing when denovoLR < the raw threshold (NucFam:1868).  The
     reference still runs CalcPostProb for such sites (main:576-587) and discards the result; nothing
     later reads it, so the restatement stops here. */
  if (par->denovo && fl[0].denovoLR < par->denovo_min_llr) {
    if (maxidx == 0) r->flags |= PM_FLAG_MONO;
    r->flags |= PM_FLAG_ROW_DROPPED;
    fill_result(c, &fl[0], r);
    r->status = PM_SITE_DENOVO_DROPPED;
    return 0;
  }
  if (maxidx == 0) { /* main:576-587 */
    if (par->denovo) { fl[0].denovo_mono = 1; CalcPostProb(c, &fl[0], 1.0); }
    else { fl[0].isMono = 1; CalcPostProb(c, &fl[0], 1 - par->theta); }
    r->flags |= PM_FLAG_MONO;
  } else {
    fl[0].isMono = 0;
    CalcPostProb(c, &fl[0], fl[0].min);
  }
  /* what the writer computes before printing: NucFam:1791 (AB) */
  if (!par->denovo && c->hdr.chr_class == PM_CHR_AUTO) CalculateAB(c, &fl[0], fl[0].min); /* NucFam:1791 */
  fl[0].denovo_mono = 0;
  fill_result(c, &fl[0], r);
  if (c->hdr.chr_class != PM_CHR_AUTO) r->ab = 0.0; /* AB keeps an older site's value there and is not printed */
  r->status = PM_SITE_EMITTED;
  for (int i = 0; i < c->nPerson; i++) {
    pm_person_result *p = &pr[i];
    int n = fl[0].tenState[i] ? 10 : 3;
    for (int g = 0; g < 10; g++) p->post[g] = g < n ? fl[0].postProb[i][g] : 0.0;
    p->dosage = fl[0].dosage[i];
    p->best = fl[0].bestGenoIdx[i];
    double pb = fl[0].postProb[i][p->best];
    int GTQual; /* NucFam:1819-1820 */
    if (pb > 0.9999999999) GTQual = 100;
    else GTQual = (int)(-10. * log10(1. - pb) + 0.5);
    p->gq = (uint8_t)(GTQual < 0 ? 0 : (GTQual > 255 ? 255 : GTQual));
    p->ten_state = fl[0].tenState[i];
  }
  return 0;
}

/* One VCF record: PedVCF::VarCallFromVCF, src/PedVCF.cpp:116-163, with FamilyLikelihoodSeq_VCF::CalcPostProb
 * (src/FamilyLikelihoodSeq_VCF.cpp:142-153) and CalcGQ (NucFam:553-569). */
int pmo_call_vcf_records(pmo_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono_in,
                         size_t n, pm_site_result *res_out, pm_person_result *person_out) {
  if (!c->par.vcf_input) { set_err("oracle ctx was not created for VCF input"); return PM_EINVAL; }
  famlk_t *k = &c->famlk[0];
  const double polyPrior = c->prior;
  for (size_t s = 0; s < n; s++) {
    pmo_load_site(c, &hdr[s], person_site + s * (size_t)c->nPerson);
    pm_site_result *r = &res_out[s];
    pm_person_result *pr = person_out + s * (size_t)c->nPerson;
    memset(r, 0, sizeof *r);
    memset(pr, 0, sizeof(*pr) * (size_t)c->nPerson);
    const int a1 = hdr[s].ref_base, a2 = hdr[s].reserved & 0xff, indel = (hdr[s].reserved >> 8) & 1;
    const double mono = mono_in[s];
    SetAlleles(k, a1, a2);                       /* PolymorphismLogLikelihood, FLSeq_VCF:85-90 */
    OptimizeFrequency(c, k);
    const double poly = -k->fmin;
    const int isTs = (a1 == 1 && a2 == 3) || (a1 == 2 && a2 == 4); /* PedVCF.cpp:23-26 */
    double llk_alt, llk_ref;
    if (!indel) {
      /* log10(polyPrior * isTs ? ts : tv): the product only selects the branch (PedVCF.cpp:143) */
      llk_alt = log10((polyPrior * isTs) != 0 ? 2.0 / (2.0 + 1) : 0.5 / (2.0 + 1)) + poly;
      llk_ref = log10(1 - polyPrior) + mono;
    } else {
      llk_alt = log10(polyPrior) + poly;         /* GetPolyPrior_indel returns the SNP prior (NucFam:313) */
      llk_ref = log10(1 - polyPrior) + mono;
    }
    double qual;
    if (llk_alt - llk_ref > 10) qual = 10.0 * (llk_alt - llk_ref);
    else {
      double posterior = 1 / (1 + pow(10, llk_ref - llk_alt));
      qual = -10 * log10(1 - posterior);
      r->var_post_prob = posterior;
    }
    /* CalcPostProb, FLSeq_VCF:142-153 */
    const double freq = k->min;
    k->isMono = 0;
    for (int i = 0; i < c->nFam; i++) {
      if (c->famSize[i] == c->famFounders[i]) {
        for (int j = 0; j < c->famFounders[i]; j++) CalcPostProb_SinglePerson(c, k, c->famFirst[i] + j, freq);
      } else if (fam_isNuclear(c, i) && c->nFam > 1 && !c->chrX && !c->chrY && !c->chrMT) { /* FLSeq_VCF:148 */
        CalcPostProb_SingleNucFam(c, k, i, freq);
      } else {
        CalcPostProb_SingleExtendedPed_BA(c, k, i, freq);
      }
    }
    r->site = (uint32_t)s; r->status = PM_SITE_EMITTED; r->maxidx = 1; r->n_hyp = 2;
    r->allele1 = (uint8_t)a1; r->allele2 = (uint8_t)a2;
    r->varllk[0] = llk_ref; r->varllk[1] = llk_alt;
    r->varllk_noprior[0] = mono; r->varllk_noprior[1] = poly;
    r->varfreq[0] = 1.0; r->varfreq[1] = freq;
    r->poly_qual = qual; r->freq = freq;
    for (int i = 0; i < c->nPerson; i++) {
      pm_person_result *p = &pr[i];
      for (int g = 0; g < 3; g++) p->post[g] = k->postProb[i][g];
      p->dosage = k->dosage[i];
      p->best = k->bestGenoIdx[i];
      double pb = k->postProb[i][p->best];
      int GTQual;
      if (pb > 0.9999999999) GTQual = 100;
      else GTQual = (int)(-10. * log10(1. - pb) + 0.5);
      p->gq = (uint8_t)(GTQual < 0 ? 0 : (GTQual > 255 ? 255 : GTQual));
    }
  }
  return PM_OK;
}

int pmo_call_glf_sites(pmo_ctx *c, const pm_site_hdr *hdr, const pm_person_site *person_site, size_t n_sites,
                       uint16_t *status_out, pm_site_result *res_out, pm_person_result *person_out) {
  for (size_t s = 0; s < n_sites; s++) {
    pmo_load_site(c, &hdr[s], person_site + s * (size_t)c->nPerson);
    pm_person_result *pr = person_out ? person_out + s * (size_t)c->nPerson : NULL;
    pm_person_result *tmp = NULL;
    if (!pr) pr = tmp = calloc((size_t)c->nPerson, sizeof *pr);
    else memset(pr, 0, sizeof(*pr) * (size_t)c->nPerson);
    int rc = call_site(c, (uint32_t)s, &res_out[s], pr);
    free(tmp);
    if (rc) return rc;
    if (status_out)
      status_out[s] = (uint16_t)(res_out[s].status | ((res_out[s].maxidx + 1) << 4) | ((res_out[s].flags & PM_FLAG_NOCALL) << 8));
  }
  return PM_OK;
}
