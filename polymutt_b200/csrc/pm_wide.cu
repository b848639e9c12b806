// pm_wide.cu — k_sites_wide: one BLOCK per site, for pedigrees made of many nuclear families and unrelated
// founders (tens to thousands of quartic units), sm_100a.
//
// What is computed: main.cpp:325-594 for one site — CalcReadStats + filters, the hypothesis likelihoods H0..H6 with
// one Brent minimisation (core/MathGold.cpp:81-177) each over  sum_f log10 L_f(p)  (FamilyLikelihoodSeq.cpp:222-240),
// CalcVarPosterior, the emit / de novo LR decisions.  How it is laid out on the GPU:
//
//   * persistent grid, one site per block at a time; the site's n_person*16 bytes are staged in shared memory by ONE
//     TMA bulk copy (cp.async.bulk + mbarrier) and the next site of the block is prefetched into L2 meanwhile;
//   * thread t owns units t, t+T, ... (U per thread); their quartic coefficients  L_u(p) = q^4 * P_u(p/q),
//     P_u(r) = B0 + B1 r + ... + B4 r^4  (pm_device.cuh) are scaled to [1,2) and stay in REGISTERS for the whole
//     Brent run.  One objective evaluation is a 4-FMA Horner step and one multiply per unit; the factor q^4 of the
//     U slots of a thread is q^(4U): a few squarings of q's mantissa, the exponent exactly in integers;
//   * the product over the block goes through warp shuffles on normalised mantissas (exponents: one REDUX) and one
//     shared-memory hop; EVERY thread then finishes the round by itself — table log10, the Brent update and the next
//     evaluation point live in its own registers, bit-identical in all threads — so a round costs ONE block barrier
//     and no broadcast (round 1 of this kernel had a one-thread serial tail between two barriers: 25 % of all stall
//     samples sat on the barrier in front of it);
//   * units whose nine conditionals are all subnormal or zero ("fragile": a family likelihood that underflows in the
//     reference) are evaluated the reference's way at every point, so that log10(0) = -inf appears exactly where
//     FamilyLikelihoodSeq.cpp:222-240 produces it;
//   * units beyond T*U (more than 4,096 units at T = 512) keep their coefficients in a global-memory scratch (L2).
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>

#include "pm_device.cuh"
#include "pm_es.cuh"
#include "pm_kernels.h"
#include "pm_site_logic.cuh"

namespace pm {

#ifdef PM_PHASE_TIMING
#define PM_TICK(slot) do { if (threadIdx.x == 0) { long long now__ = clock64(); s_ws.phase[slot] += (unsigned long long)(now__ - s_ws.t_last); s_ws.t_last = now__; } } while (0)
#else
#define PM_TICK(slot) do { } while (0)
#endif

// Hot tables in STATIC shared memory: their addresses are link-time constants, so a look-up lut[byte] needs no base
// register and no address arithmetic beyond the scaled index (the out-of-line helpers name them directly).
__shared__ __align__(16) double s_lut[256];      // 10^(-i/10), host-computed
__shared__ __align__(16) double s_mut[100];      // genotype mutation matrix
__shared__ __align__(16) double s_log_inv[128];  // table-driven log10 of a mantissa in [1,2)
__shared__ __align__(16) double s_log_tab[128];

constexpr int kMaxSpec = 12;  // 10 points at the default --prec 1e-4

struct WideShared {
  pm_site_result r;                   // written by thread 0 only
  double red_m[32];                   // per-warp partial products of the current round (slots >= #warps stay at 1)
  int red_e[32];
  double red0_m[32];                  // the H0 product (first round of H1 under --denovo)
  int red0_e[32];
  int4 red_i[2][32];                  // per-warp integer partials (depth, samples, mapq, lk sum), double-buffered by site parity
  double next[4];                     // broadcast by thread 0 after every round: p, r = p/q, q^(4 * units per partial product)
  // The monotone path (see spec_table_ol): evaluation points Brent visits when no point beats the first one
  double spec_p[kMaxSpec], spec_r[kMaxSpec], spec_q[kMaxSpec];  // p, p/q, q^(4 * units per speculative partial product)
  double spec_qn[9][kMaxSpec];        // q^(4 n), n = 0..8 units of a thread (the one-pass H1..H3 evaluation)
  double spec_m[kMaxSpec][17];        // per point, per warp: partial products of the speculative round (17: bank padding)
  int spec_e[kMaxSpec][17];
  double spec_ll[kMaxSpec];           // log10 likelihood at every point of the path
  int spec_n, spec_complete;          // points on the path (<= kMaxSpec); 1 = the path ends with Brent's stopping rule
  BrentState brent;                   // driven by thread 0 (brent_round_ol)
  double h0;                          // log10 of the H0 product
  double lk_mono;                     // MonomorphismLogLikelihood(refBase) of the current site
  unsigned n_eval;                    // rounds of the current chain
  int more;                           // another round?
  int ibcast[4];
  unsigned long long mbar;            // mbarrier of the site buffer (TMA bulk copies)
  unsigned long long phase[8];        // PM_PHASE_TIMING: cycles per phase (thread 0)
  long long t_last;
};
// Static shared memory like the tables: no generic-to-shared address arithmetic on the serial path of a round.
__shared__ __align__(16) WideShared s_ws;

// bar.sync 0 by hand: warp 0 and the other warps reach the per-round barriers from two copies of the round loop
// (different program counters, same count), which the hardware barrier allows and __syncthreads() does not promise.
__device__ __forceinline__ void block_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void tma_issue_site(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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

// ---- (mantissa in [1,2), integer exponent) products ------------------------------------------------
// A value of exactly 0 is kept as mantissa 0: log10 -> -inf, which is what the reference computes for a family
// likelihood that underflowed (FamilyLikelihoodSeq.cpp:222-240).
struct ME { double m; int e; };
__device__ __forceinline__ ME me_split(double x) {  // x >= 0, finite, possibly subnormal or zero (rare paths)
  ME r;
  int hi = __double2hiint(x);
  int ex = (hi >> 20) & 0x7ff;
  r.e = 0;
  if (ex == 0) {
    if (!(x > 0.0)) { r.m = 0.0; return r; }
    x *= 1.3407807929942597e154;  // 2^512
    r.e = -512;
    hi = __double2hiint(x);
    ex = (hi >> 20) & 0x7ff;
  }
  r.e += ex - 1023;
  r.m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  return r;
}
// x is a normal positive number (products of in-range factors); zero is handled by the caller
__device__ __forceinline__ ME me_split_pos(double x) {
  ME r;
  const int hi = __double2hiint(x);
  r.e = ((hi >> 20) & 0x7ff) - 1023;
  r.m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  return r;
}
__device__ __forceinline__ void me_mul(ME &a, double x) {  // a *= x for an arbitrary x >= 0 (rare paths)
  const ME b = me_split(x);
  const double m = a.m * b.m;
  if (!(m > 0.0)) { a.m = 0.0; return; }
  const ME c = me_split_pos(m);
  a.m = c.m; a.e += b.e + c.e;
}

// 1/x to the last bit or so (MUFU.RCP64H + two Newton steps, ~50 cycles against ~125 for the IEEE division):
// on the serial path of every Brent round; x is a normal number.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// log10(m * 2^e) for a mantissa m in [1,2) (or 0 -> -inf).  m = c (1 + r) with c from a 128-entry table,
// |r| <= 2^-8, log1p(r) by its series to r^7 (truncation < 2^-67); absolute error ~1e-16, far below the
// rounding noise of the reference's own sum of per-family log10 values.
__device__ __forceinline__ double log10_me(double m, int e) {
  if (!(m > 0.0)) return -CUDART_INF;
  const int i = (__double2hiint(m) >> 13) & 127;
  const double r = fma(m, s_log_inv[i], -1.0);
  const double r2 = r * r;
  const double a = fma(1.0 / 3.0, r, -0.5), b = fma(1.0 / 5.0, r, -0.25), c = fma(1.0 / 7.0, r, -1.0 / 6.0);
  const double r4 = r2 * r2;
  const double q = fma(c, r4, fma(b, r2, a));
  const double l1p = fma(r2, q, r);
  return fma((double)e, kLog10_2, fma(l1p, 0.43429448190325182765, s_log_tab[i]));
}

// Brent (core/MathGold.cpp:81-177) exactly as pm_device.cuh's brent_feed, the one division of the parabolic step
// taken by fast_rcp (the step differs from the IEEE quotient in the last bit at most).
__device__ __forceinline__ bool brent_feed_fast(BrentState &s, double fu, double tol) {
  if (s.first) {
    s.fmin = s.fw = s.fv = fu;
    s.first = 0;
  } else {
    const double u = s.u;
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
  if (s.iter > PM_ITMAX) return false;
  const double middle = 0.5 * (s.a + s.c);
  const double tol1 = tol * fabs(s.min) + PM_ZEPS;
  const double tol2 = 2.0 * tol1;
  if (fabs(s.min - middle) <= (tol2 - 0.5 * (s.c - s.a))) return false;
  const double golden_delta = s.min >= middle ? s.a - s.min : s.c - s.min;
  if (fabs(s.delta) > tol1) {
    const double r = (s.min - s.w) * (s.fmin - s.fv);
    double q = (s.min - s.v) * (s.fmin - s.fw);
    double p = (s.min - s.v) * q - (s.min - s.w) * r;
    q = 2.0 * (q - r);
    if (q > 0.0) p = -p;
    q = fabs(q);
    const double temp = s.delta;
    s.delta = s.d;
    if (fabs(p) >= fabs(0.5 * q * temp) || p <= q * (s.a - s.min) || p >= q * (s.c - s.min)) {
      s.delta = golden_delta;
      s.d = PM_CGOLD * s.delta;
    } else {
      s.d = p * fast_rcp(q);
      const double u = s.min + s.d;
      if (u - s.a < tol2 || s.c - u < tol2) s.d = sign_d(tol1, middle - s.min);
    }
  } else {
    s.delta = golden_delta;
    s.d = PM_CGOLD * s.delta;
  }
  s.u = fabs(s.d) >= tol1 ? s.min + s.d : s.min + sign_d(tol1, s.d);
  return true;
}

// ---- quartic coefficients of one unit, out of line ---------------------------------------------------
// Their register pressure (ten look-ups, three 10-term dot products) stays out of the kernel's hot loop allocation.
// N autosomal nuclear families with the same number of kids, in lockstep (the chrX / chrY / MT rules live in the NA
// instance only): every step is written for all N before the next one, so that N independent dependency chains are in
// flight — the set-up is latency-bound, not throughput-bound.  Same value as quartic_from_conditionals(
// unit_conditionals(...)) up to the last bit: the nine conditionals are not formed one by one, the pairs that share a
// kid product are summed first (19 multiply-adds instead of 25).  The rows of the mutation matrix are read once for all N.
template <int N>
__device__ __forceinline__ void nuclear_quartic_n(const uint4 *recs, const int (&first)[N], int nkids, int g11, int g12, int g22, int denovo,
                                                  double (&b)[N][5]) {
  double f11[N], f12[N], f22[N], m11[N], m12[N], m22[N];
  double p0[N], p1[N], p2[N], p4[N], p5[N], p8[N];  // kid products: p1 carries 2^nkids, p4 4^nkids, p5 2^nkids
#pragma unroll
  for (int n = 0; n < N; n++) {
    const uint8_t *rb = reinterpret_cast<const uint8_t *>(recs + first[n]);
    f11[n] = s_lut[rb[g11]]; f12[n] = s_lut[rb[g12]]; f22[n] = s_lut[rb[g22]];
    m11[n] = s_lut[rb[16 + g11]]; m12[n] = s_lut[rb[16 + g12]]; m22[n] = s_lut[rb[16 + g22]];
    p0[n] = p1[n] = p2[n] = p4[n] = p5[n] = p8[n] = 1.0;
  }
  for (int k = 0; k < nkids; k++) {
    double d11[N], d12[N], d22[N];
    if (denovo) {  // CalcDenovoMutLk, NucFam:1553-1562
      double l[N][10];
#pragma unroll
      for (int n = 0; n < N; n++) {
        const uint4 rk = recs[first[n] + 2 + k];
        l[n][0] = s_lut[__byte_perm(rk.x, 0, 0x4440)]; l[n][1] = s_lut[__byte_perm(rk.x, 0, 0x4441)];
        l[n][2] = s_lut[__byte_perm(rk.x, 0, 0x4442)]; l[n][3] = s_lut[__byte_perm(rk.x, 0, 0x4443)];
        l[n][4] = s_lut[__byte_perm(rk.y, 0, 0x4440)]; l[n][5] = s_lut[__byte_perm(rk.y, 0, 0x4441)];
        l[n][6] = s_lut[__byte_perm(rk.y, 0, 0x4442)]; l[n][7] = s_lut[__byte_perm(rk.y, 0, 0x4443)];
        l[n][8] = s_lut[__byte_perm(rk.z, 0, 0x4440)]; l[n][9] = s_lut[__byte_perm(rk.z, 0, 0x4441)];
      }
      const double2 *r11 = reinterpret_cast<const double2 *>(s_mut + g11 * 10);
      const double2 *r12 = reinterpret_cast<const double2 *>(s_mut + g12 * 10);
      const double2 *r22 = reinterpret_cast<const double2 *>(s_mut + g22 * 10);
      double a11[N], b11[N], a12[N], b12[N], a22[N], b22[N];
#pragma unroll
      for (int n = 0; n < N; n++) a11[n] = b11[n] = a12[n] = b12[n] = a22[n] = b22[n] = 0.0;
#pragma unroll
      for (int g = 0; g < 5; g++) {
        const double2 x = r11[g], y = r12[g], z = r22[g];
#pragma unroll
        for (int n = 0; n < N; n++) {
          a11[n] = fma(x.x, l[n][2 * g], a11[n]); b11[n] = fma(x.y, l[n][2 * g + 1], b11[n]);
          a12[n] = fma(y.x, l[n][2 * g], a12[n]); b12[n] = fma(y.y, l[n][2 * g + 1], b12[n]);
          a22[n] = fma(z.x, l[n][2 * g], a22[n]); b22[n] = fma(z.y, l[n][2 * g + 1], b22[n]);
        }
      }
#pragma unroll
      for (int n = 0; n < N; n++) { d11[n] = a11[n] + b11[n]; d12[n] = a12[n] + b12[n]; d22[n] = a22[n] + b22[n]; }
    } else {
#pragma unroll
      for (int n = 0; n < N; n++) {
        const uint8_t *kb = reinterpret_cast<const uint8_t *>(recs + first[n] + 2 + k);
        d11[n] = s_lut[kb[g11]]; d12[n] = s_lut[kb[g12]]; d22[n] = s_lut[kb[g22]];
      }
    }
    // likelihoodONEKid{,_denovo}, NucFam:1202-1296 (autosome), the halves and quarters taken out
#pragma unroll
    for (int n = 0; n < N; n++) {
      p0[n] *= d11[n];
      p1[n] *= d11[n] + d12[n];
      p2[n] *= d12[n];
      p4[n] *= fma(2.0, d12[n], d11[n]) + d22[n];
      p5[n] *= d12[n] + d22[n];
      p8[n] *= d22[n];
    }
  }
  // exact powers of two: 2 * (1/2)^nkids for B3 and B1, 4 * (1/4)^nkids for the double-heterozygote term
  const double h1 = __hiloint2double((1023 + 1 - nkids) << 20, 0), h4 = __hiloint2double((1023 + 2 - 2 * nkids) << 20, 0);
#pragma unroll
  for (int n = 0; n < N; n++) {
    b[n][4] = p0[n] * (f11[n] * m11[n]);
    b[n][3] = (p1[n] * h1) * fma(f11[n], m12[n], f12[n] * m11[n]);
    b[n][2] = fma(p4[n] * h4, f12[n] * m12[n], p2[n] * fma(f11[n], m22[n], f22[n] * m11[n]));
    b[n][1] = (p5[n] * h1) * fma(f12[n], m22[n], f22[n] * m12[n]);
    b[n][0] = p8[n] * (f22[n] * m22[n]);
  }
}

// Scaled coefficients of one slot into `out` (5 doubles).  Returns the power of two taken out; kSlotNeutral: the
// slot holds no unit; kSlotFragile: a fragile unit (see the file header), the slot gets the neutral coefficients.
constexpr int kSlotNeutral = 1 << 20, kSlotFragile = 1 << 21;
__device__ __forceinline__ int scale_slot(const double (&b)[5], bool nuclear, double *out) {
  // all coefficients are >= 0: the exponent of their sum is within 3 of the largest one's
  const double sum = ((b[0] + b[1]) + (b[2] + b[3])) + b[4];
  const int ex = (__double2hiint(sum) >> 20) & 0x7ff;
  // A family whose coefficients are all below 2^-957 may underflow (to exactly 0, or into the subnormals) in the
  // reference's  sum_j C_j * prior_j(p)  with prior_j >= 2^-54: it is evaluated the reference's way.  Above that the
  // reference's value is a normal number and the scaled form is exact to the last bits.
  if (ex < 66 && nuclear) {
    out[0] = 0.25; out[1] = 1.0; out[2] = 1.5; out[3] = 1.0; out[4] = 0.25;
    return kSlotFragile;
  }
  // Scaled so that the coefficients sum to [1,2): with r = p/q in [1e-4, 1e4] the value P'(r) q^4 = L'(p) lies in
  // [2^-57, 2] and P'(r) in [2^-57, 2^55]: products of eight units need no exponent handling.
  const double sc = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023-ex), exact (ex = 0: an all-zero single founder, P = 0)
  out[0] = b[0] * sc; out[1] = b[1] * sc; out[2] = b[2] * sc; out[3] = b[3] * sc; out[4] = b[4] * sc;
  return (ex - 1023) & 0xfffff;
}
// The same without the fragile branch (the one-pass evaluation gives a site with a fragile unit up anyway): the scaled
// coefficients are always written, the code carries kSlotFragile where scale_slot would have returned it.
__device__ __forceinline__ int scale_slot_nb(const double (&b)[5], bool nuclear, double (&out)[5]) {
  const double sum = ((b[0] + b[1]) + (b[2] + b[3])) + b[4];
  const int ex = (__double2hiint(sum) >> 20) & 0x7ff;
  const double sc = __hiloint2double((2046 - ex) << 20, 0);
  out[0] = b[0] * sc; out[1] = b[1] * sc; out[2] = b[2] * sc; out[3] = b[3] * sc; out[4] = b[4] * sc;
  return ((ex - 1023) & 0xfffff) | ((ex < 66 && nuclear) ? kSlotFragile : 0);
}
__device__ __forceinline__ int one_slot(const uint4 *recs, int d, int g11, int g12, int g22, int mode, double *out);

// Two slots of a thread per call (slot = packed unit descriptor, -1 = none): when both are nuclear families with the
// same number of kids they go through the lockstep code above.  out0 / out1: 5 doubles each (the caller's staging
// array).  Returns the two slots' codes (20 bits + flags each) in an int2.
#ifdef PM_INLINE_SLOTS
#define PM_SLOT_INLINE __forceinline__
#else
#define PM_SLOT_INLINE __noinline__
#endif
__device__ PM_SLOT_INLINE int2 slot_pair_ol(const uint4 *recs, int d0, int d1, int g11, int g12, int g22, int mode, double *out0, double *out1) {
  int2 ret;
  const int nk0 = ((d0 >> 20) & 0xff) - 1, nk1 = ((d1 >> 20) & 0xff) - 1;
  if (d0 >= 0 && d1 >= 0 && nk0 == nk1 && nk0 >= 1) {
    const int first[2] = {d0 & 0xfffff, d1 & 0xfffff};
    double b[2][5];
    nuclear_quartic_n<2>(recs, first, nk0, g11, g12, g22, mode & 1, b);
    ret.x = scale_slot(b[0], true, out0);
    ret.y = scale_slot(b[1], true, out1);
    return ret;
  }
  ret.x = one_slot(recs, d0, g11, g12, g22, mode, out0);
  ret.y = one_slot(recs, d1, g11, g12, g22, mode, out1);
  return ret;
}
__device__ __forceinline__ int one_slot(const uint4 *recs, int d, int g11, int g12, int g22, int mode, double *out) {
  if (d < 0) { out[0] = 0.25; out[1] = 1.0; out[2] = 1.5; out[3] = 1.0; out[4] = 0.25; return kSlotNeutral; }
  const int first = d & 0xfffff, nkids = ((d >> 20) & 0xff) - 1;
  double b[1][5];
  if (nkids < 0) {  // lkSinglePerson, NucFam:987-1004, times (p+q)^2
    const uint8_t *rb = reinterpret_cast<const uint8_t *>(recs + first);
    const double l11 = s_lut[rb[g11]], l12 = s_lut[rb[g12]], l22 = s_lut[rb[g22]];
    b[0][4] = l11; b[0][3] = 2.0 * (l11 + l12); b[0][2] = l11 + 4.0 * l12 + l22; b[0][1] = 2.0 * (l12 + l22); b[0][0] = l22;
  } else {
    const int f1[1] = {first};
    nuclear_quartic_n<1>(recs, f1, nkids, g11, g12, g22, mode & 1, b);
  }
  return scale_slot(b[0], nkids >= 0, out);
}
struct PairOut { double b[2][5]; int c0, c1; };
__device__ __noinline__ PairOut slot_pair_val_ol(const uint4 *recs, int d0, int d1, int g11, int g12, int g22, int mode) {
  PairOut o;
  const int2 c = slot_pair_ol(recs, d0, d1, g11, g12, g22, mode, o.b[0], o.b[1]);
  o.c0 = c.x; o.c1 = c.y;
  return o;
}
__device__ PM_SLOT_INLINE int slot_single_ol(const uint4 *recs, int d, int g11, int g12, int g22, int mode, double *out) {
  return one_slot(recs, d, g11, g12, g22, mode, out);
}

// One unit, the three hypotheses H1..H3 = (ref, ts), (ref, tv1), (ref, tv2) of a site at once (autosome).  They share the
// reference-homozygous genotype and, above all, every byte extraction and table look-up of the unit's records: the
// ten likelihoods of a kid are fetched once, seven rows of the mutation matrix (ref/ref, three ref/alt, three alt/alt)
// give the mutation-mixed likelihoods of all three hypotheses — 70 multiply-adds and 11 look-ups per kid instead of
// 90 and 30 — and the seventeen independent dependency chains keep the FP64 pipe busy where one hypothesis alone
// waits on latency.  b[h][0..4]: the unscaled coefficients (first column, nkids; -1 = a single founder); bit for bit
// what nuclear_quartic_n / one_slot give for each hypothesis on its own.
__device__ __forceinline__ void unit_h123(const uint4 *recs, int first, int nkids, int ref, int denovo, double (&b)[3][5]) {
  const int alt[3] = {poly_ts(ref), poly_tvs1(ref), poly_tvs2(ref)};
  const int grr = geno_index(ref, ref);
  int gra[3], gaa[3];
#pragma unroll
  for (int h = 0; h < 3; h++) { gra[h] = geno_index(ref, alt[h]); gaa[h] = geno_index(alt[h], alt[h]); }
  const uint8_t *rb = reinterpret_cast<const uint8_t *>(recs + first);
  if (nkids < 0) {  // lkSinglePerson, NucFam:987-1004, times (p+q)^2
    const double l11 = s_lut[rb[grr]];
#pragma unroll
    for (int h = 0; h < 3; h++) {
      const double l12 = s_lut[rb[gra[h]]], l22 = s_lut[rb[gaa[h]]];
      b[h][4] = l11; b[h][3] = 2.0 * (l11 + l12); b[h][2] = l11 + 4.0 * l12 + l22; b[h][1] = 2.0 * (l12 + l22); b[h][0] = l22;
    }
    return;
  }
  // the mutation-mixed (or plain) likelihoods of kid k under the seven genotypes the three hypotheses name
  auto kid = [&](int k, double &drr, double (&dra)[3], double (&daa)[3]) {
    if (denovo) {  // CalcDenovoMutLk, NucFam:1553-1562
      const uint4 rk = recs[first + 2 + k];
      double l[10];
      l[0] = s_lut[__byte_perm(rk.x, 0, 0x4440)]; l[1] = s_lut[__byte_perm(rk.x, 0, 0x4441)];
      l[2] = s_lut[__byte_perm(rk.x, 0, 0x4442)]; l[3] = s_lut[__byte_perm(rk.x, 0, 0x4443)];
      l[4] = s_lut[__byte_perm(rk.y, 0, 0x4440)]; l[5] = s_lut[__byte_perm(rk.y, 0, 0x4441)];
      l[6] = s_lut[__byte_perm(rk.y, 0, 0x4442)]; l[7] = s_lut[__byte_perm(rk.y, 0, 0x4443)];
      l[8] = s_lut[__byte_perm(rk.z, 0, 0x4440)]; l[9] = s_lut[__byte_perm(rk.z, 0, 0x4441)];
      auto row = [&](int g) {
        const double2 *r = reinterpret_cast<const double2 *>(s_mut + g * 10);
        double a = 0.0, c = 0.0;
#pragma unroll
        for (int i = 0; i < 5; i++) { const double2 x = r[i]; a = fma(x.x, l[2 * i], a); c = fma(x.y, l[2 * i + 1], c); }
        return a + c;
      };
      drr = row(grr);
#pragma unroll
      for (int h = 0; h < 3; h++) { dra[h] = row(gra[h]); daa[h] = row(gaa[h]); }
    } else {
      const uint8_t *kb = reinterpret_cast<const uint8_t *>(recs + first + 2 + k);
      drr = s_lut[kb[grr]];
#pragma unroll
      for (int h = 0; h < 3; h++) { dra[h] = s_lut[kb[gra[h]]]; daa[h] = s_lut[kb[gaa[h]]]; }
    }
  };
  // kid products (p1 carries 2^nkids, p4 4^nkids, p5 2^nkids); likelihoodONEKid{,_denovo}, NucFam:1202-1296 (autosome),
  // the halves and quarters taken out.  The first kid starts them (1.0 * x = x exactly: same bits as a loop from 1.0).
  double p0 = 1.0, p1[3], p2[3], p4[3], p5[3], p8[3];
#pragma unroll
  for (int h = 0; h < 3; h++) p1[h] = p2[h] = p4[h] = p5[h] = p8[h] = 1.0;
  if (nkids > 0) {
    double drr, dra[3], daa[3];
    kid(0, drr, dra, daa);
    p0 = drr;
#pragma unroll
    for (int h = 0; h < 3; h++) {
      p1[h] = drr + dra[h];
      p2[h] = dra[h];
      p4[h] = fma(2.0, dra[h], drr) + daa[h];
      p5[h] = dra[h] + daa[h];
      p8[h] = daa[h];
    }
  }
  for (int k = 1; k < nkids; k++) {
    double drr, dra[3], daa[3];
    kid(k, drr, dra, daa);
    p0 *= drr;
#pragma unroll
    for (int h = 0; h < 3; h++) {
      p1[h] *= drr + dra[h];
      p2[h] *= dra[h];
      p4[h] *= fma(2.0, dra[h], drr) + daa[h];
      p5[h] *= dra[h] + daa[h];
      p8[h] *= daa[h];
    }
  }
  // the parents' look-ups come after the kid loop: fourteen values less to keep alive across it
  const double frr = s_lut[rb[grr]], mrr = s_lut[rb[16 + grr]];
  const double h1 = __hiloint2double((1023 + 1 - nkids) << 20, 0), h4 = __hiloint2double((1023 + 2 - 2 * nkids) << 20, 0);
  const double b4 = p0 * (frr * mrr);
#pragma unroll
  for (int h = 0; h < 3; h++) {
    const double fra = s_lut[rb[gra[h]]], faa = s_lut[rb[gaa[h]]], mra = s_lut[rb[16 + gra[h]]], maa = s_lut[rb[16 + gaa[h]]];
    b[h][4] = b4;
    b[h][3] = (p1[h] * h1) * fma(frr, mra, fra * mrr);
    b[h][2] = fma(p4[h] * h4, fra * mra, p2[h] * fma(frr, maa, faa * mrr));
    b[h][1] = (p5[h] * h1) * fma(fra, maa, faa * mra);
    b[h][0] = p8[h] * (faa * maa);
  }
}
// chrX / chrY / MT slot; mode = denovo | chr_class << 1
__device__ __noinline__ int slot_na_ol(const uint4 *recs, int d, int g11, int g12, int g22, int mode, double *out) {
  if (d < 0) { out[0] = 0.25; out[1] = 1.0; out[2] = 1.5; out[3] = 1.0; out[4] = 0.25; return kSlotNeutral; }
  double B[5];
  DevUnit u;
  u.first = d & 0xfffff; u.nkids = ((d >> 20) & 0xff) - 1; u.sex = (d >> 28) & 3; u.kid0 = 0;
  unit_quartic_nonauto(recs, u, g11, g12, g22, (mode & 1) != 0, (mode >> 1) & 3, 0, s_lut, s_mut, B);
  return scale_slot(B, u.nkids >= 0, out);
}

// A fragile nuclear unit at frequency p, the reference's way: sum_j parentConditional[j] * parentPrior[j]
// (NucFam:941-985, 1041-1132), every product and sum rounded on its own (the host code has no FMA contraction), so that
// the value — and whether it is exactly 0 — is the reference's.
__device__ __noinline__ double unit_ref_likelihood_ol(const uint4 *recs, int first, int nkids, int g11, int g12, int g22, int mode, double p) {
  const bool denovo = (mode & 1) != 0;
  const int cls = (mode >> 1) & 3;
  double C[9], pp[9];
  if (cls == PM_CHR_AUTO || denovo) {
    unit_conditionals(recs, first, nkids, g11, g12, g22, denovo, s_lut, s_mut, C);
    parent_priors(p, pp);
  } else {
    unit_conditionals_nonauto(recs, first, nkids, g11, g12, g22, cls, 0, s_lut, C);
    parent_priors_nonauto(cls, p, pp);
  }
  double sum = 0.0;
  for (int j = 0; j < 9; j++) sum = __dadd_rn(sum, __dmul_rn(C[j], pp[j]));
  return sum;
}

// ES instances: extended families take part in every evaluation through a thread-serial Elston-Stewart peel, family e
// on thread e of the block (FLSeq.cpp:222-240 sums log10 over all families; here: one more factor of the product).
// A free function on purpose, and only ever named under `if constexpr (ES)` (a member function let the
// evaluator's shared-memory pointers lose their address space).
__device__ __noinline__ double es_factor(const DevRun *run, const uint4 *recs, int cls, int e, int g11, int g12, int g22, bool denovo, double p) {
  const DevFam f = run->fams[run->es_fams[e]];
  return denovo ? es_likelihood<10>(run, f, recs, g11, g12, g22, true, p, s_lut, s_mut, -1, -1, cls)
                : es_likelihood<3>(run, f, recs, g11, g12, g22, false, p, s_lut, s_mut, -1, -1, cls);
}

// packed unit descriptor kept in a register for the whole kernel: first | (nkids + 1) << 20 | sex << 28; -1 = no unit
__device__ __forceinline__ int desc_pack(const DevUnit &u) { return u.first | ((u.nkids + 1) << 20) | (u.sex << 28); }
__device__ __forceinline__ int desc_first(int d) { return d & 0xfffff; }
__device__ __forceinline__ int desc_nkids(int d) { return ((d >> 20) & 0xff) - 1; }
__device__ __forceinline__ int desc_sex(int d) { return (d >> 28) & 3; }

// Everything of a thread's share of the objective that is not a register-resident quartic, at one evaluation point, as
// one product (mantissa, exponent): fragile units the reference's way, the units beyond T*U from the L2 scratch,
// extended families by their peel.  One out-of-line call per point keeps these rare or slow paths — and their
// registers — out of the evaluation loops.  h0: the factors of the hom-ref hypothesis instead (likelihood at p = 1).
struct RareCtx {
  const DevRun *run;
  const uint4 *recs;
  const double *spill;
  unsigned fragile, spill_fragile;
  int U, T, t, n_units, g11, g12, g22, mode, cls, denovo;
  int desc[8];
};
template <bool ES>
__device__ __noinline__ ME rare_factor_ol(const RareCtx *c, double p, double r, int h0) {
  ME x;
  x.m = 1.0; x.e = 0;
  const double pe = h0 ? 1.0 : p;
  for (int k = 0; k < c->U; k++)
    if (c->fragile >> k & 1) me_mul(x, unit_ref_likelihood_ol(c->recs, desc_first(c->desc[k]), desc_nkids(c->desc[k]), c->g11, c->g12, c->g22, c->mode, pe));
  if (c->spill) {
    const double q = 1.0 - p;
    const double q4 = (q * q) * (q * q);
    int j = 0;
    for (int u = c->T * c->U + c->t; u < c->n_units; u += c->T, j++) {
      const double *co = c->spill + (size_t)(u - c->T * c->U) * 5;
      if (c->spill_fragile >> j & 1) {
        const DevUnit du = c->run->units[u];
        me_mul(x, unit_ref_likelihood_ol(c->recs, du.first, du.nkids, c->g11, c->g12, c->g22, c->mode, pe));
        if (!h0) me_mul(x, 0.25);  // Kn counted the slot's neutral (1 + r)^4 q^4 / 4, which is not evaluated here
      } else if (h0) {
        me_mul(x, co[4]);
      } else {
        me_mul(x, fma(fma(fma(fma(co[4], r, co[3]), r, co[2]), r, co[1]), r, co[0]));
        me_mul(x, q4);
      }
    }
  }
  if constexpr (ES)  // (h0: a founder's prior at p = 1 is (1, 0, 0) whatever the second allele: H1's alleles give H0's value)
    for (int f = c->t; f < c->run->n_es; f += c->T)
      me_mul(x, es_factor(c->run, c->recs, c->cls, f, c->g11, c->g12, c->g22, h0 ? true : c->denovo != 0, pe));
  return x;
}

__device__ __noinline__ void var_posterior_ol(pm_site_result *r, int ref, int n) { var_posterior(*r, ref, n); }
// The same by a whole warp (all 32 lanes call it; *r is final and visible to them): lane i takes the exp10 of hypothesis
// i, the sum is formed in the reference's order — same bits as var_posterior, a third of its dependent latency.  Lane 0
// writes the result.
__device__ __noinline__ void var_posterior_warp_ol(pm_site_result *r, int ref, int n) {
  const int lane = threadIdx.x & 31;
  int maxidx = 0;
  double mx = r->varllk[0];
  for (int i = 0; i < n; i++) if (mx < r->varllk[i]) { mx = r->varllk[i]; maxidx = i; }
  // (on a monomorphic site of a large pedigree every other hypothesis lies hundreds of log10 units below the best one:
  // 10^x is exactly 0 below x = -400 and exactly 1 at x = 0, so the warp usually skips exp10 altogether — same bits)
  const double x = lane < n ? r->varllk[lane] - mx : -1000.0;
  const bool trivial = x == 0.0 || x < -400.0;
  double e = x == 0.0 ? 1.0 : 0.0;
  if (!__all_sync(0xffffffffu, trivial)) e = lane < n ? exp10(x) : 0.0;
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += __shfl_sync(0xffffffffu, e, i);
  if (lane == 0) {
    r->var_post_prob = sum == 1.0 ? 1.0 : 1 / sum;
    int a1, a2;
    if (maxidx == 0) {
      int idx = 1;
      double m2 = r->varllk[1];
      for (int i = 1; i < 4; i++) if (m2 < r->varllk[i]) { m2 = r->varllk[i]; idx = i; }
      hyp_alleles(idx, ref, a1, a2);
    } else {
      hyp_alleles(maxidx, ref, a1, a2);
    }
    r->allele1 = (uint8_t)a1; r->allele2 = (uint8_t)a2;
    r->maxidx = (int8_t)maxidx;
    r->n_hyp = (uint8_t)n;
    r->poly_qual = (r->var_post_prob > 0.9999999999) ? 100.0 : -10 * log10(1 - r->var_post_prob);
  }
  __syncwarp();
}
__device__ __noinline__ int site_decide_ol(const DevRun *run, pm_site_result *r, double lk_mono) { return site_decide(run, *r, lk_mono) ? 1 : 0; }

__device__ __forceinline__ void publish_next(double p, int log2per) {
  const double q = 1.0 - p;
  double qa = q * q;
  qa *= qa;
  for (int i = 0; i < log2per; i++) qa *= qa;
  s_ws.next[0] = p; s_ws.next[1] = p * fast_rcp(q); s_ws.next[2] = qa;
}

// The serial part of a Brent round, thread 0 only, out of line (its registers — the whole Brent state — stay out of the
// evaluation loop's allocation; the state lives in static shared memory).  Reads the per-warp partial products, takes
// log10, feeds Brent (core/MathGold.cpp:81-177), and publishes the next evaluation point: p, r = p/q and
// q^(4 * 2^log2per), the factor every partial product of 2^log2per units starts from.
__device__ __noinline__ void brent_round_ol(int nwarp, double tol, int log2per) {
  // <= 16 warps, every partial mantissa in [1,2) (or 0): their product needs no intermediate renormalisation;
  // slots beyond nwarp hold 1 (set once per kernel), so four are always read
  double m0 = 1.0, m1 = 1.0;
  int e = 0;
  for (int w = 0; w < nwarp; w += 4) {
    const double2 a = *reinterpret_cast<const double2 *>(&s_ws.red_m[w]), b = *reinterpret_cast<const double2 *>(&s_ws.red_m[w + 2]);
    const int4 ee = *reinterpret_cast<const int4 *>(&s_ws.red_e[w]);
    m0 *= a.x * a.y; m1 *= b.x * b.y;
    e += (ee.x + ee.y) + (ee.z + ee.w);
  }
  const double mm = m0 * m1;
  const ME a = me_split_pos(mm);
  const double ll = mm > 0.0 ? log10_me(a.m, a.e + e) : -CUDART_INF;
  BrentState st = s_ws.brent;
  const bool more = brent_feed_fast(st, -ll, tol);
  s_ws.brent = st;
  s_ws.n_eval++;
  s_ws.more = more ? 1 : 0;
  publish_next(st.u, log2per);
}
// The monotone path.  OptimizeFrequency (NucFam:432-444) enters Brent with a = 1e-4, c = 0.5 and the first point
// b = 0.9999 OUTSIDE [a, c].  As long as no later point beats f(b) — the objective of a monomorphic site falls all the
// way to p = 0.9999 — every new point becomes the bound `a`, c stays 0.5 < a, the acceptance window of the parabolic
// step, (q (a - min), q (c - min)), is empty, and every step is the golden-section one: the points depend on (a, c,
// min, tol) only, not on the data.  At --prec 1e-4 that is 10 points, ended by Brent's own stopping rule.  This runs
// the real state machine once per kernel on a stand-in objective (first value 0, every later one 1) and records them.
__device__ __noinline__ void spec_table_ol(double tol, int log2sper) {
  BrentState st;
  brent_begin(st);
  int n = 0, complete = 0;
  while (n < kMaxSpec) {
    const double p = st.u, q = 1.0 - p;
    double qa = q * q;
    qa *= qa;
    for (int i = 0; i < log2sper; i++) qa *= qa;
    s_ws.spec_p[n] = p; s_ws.spec_r[n] = p * fast_rcp(q); s_ws.spec_q[n] = qa;
    {
      const double q4 = (q * q) * (q * q);
      double qn = 1.0;
      for (int i = 0; i <= 8; i++) { s_ws.spec_qn[i][n] = qn; qn *= q4; }
    }
    n++;
    if (!brent_feed_fast(st, n == 1 ? 0.0 : 1.0, tol)) { complete = 1; break; }
  }
  s_ws.spec_n = n; s_ws.spec_complete = complete;
  // the slots behind the last point repeat it: the one-pass evaluation works on a fixed number of points without selects
  for (int j = n; j < kMaxSpec; j++) {
    s_ws.spec_p[j] = s_ws.spec_p[n - 1]; s_ws.spec_r[j] = s_ws.spec_r[n - 1]; s_ws.spec_q[j] = s_ws.spec_q[n - 1];
    for (int i = 0; i <= 8; i++) s_ws.spec_qn[i][j] = s_ws.spec_qn[i][n - 1];
  }
}

// After the speculative round (thread 0): the likelihoods at all points of the monotone path are in spec_ll.  If none
// of them reaches the first one the path is the one Brent takes and its result is the first point; otherwise Brent is
// replayed on the recorded values up to the point where it leaves the path, and goes on one evaluation per round.
__device__ __noinline__ void spec_resolve_ol(double tol, int log2per) {
  const int n = s_ws.spec_n;
  const double ll0 = s_ws.spec_ll[0];
  bool on_path = s_ws.spec_complete != 0;
  for (int k = 1; k < n; k++) on_path = on_path && (s_ws.spec_ll[k] < ll0);  // f_k > f_0  (a NaN leaves the path)
  if (on_path) {
    BrentState st;
    brent_begin(st);
    st.first = 0; st.fmin = -ll0; st.iter = n;
    s_ws.brent = st;
    s_ws.n_eval = (unsigned)n;
    s_ws.more = 0;
    return;
  }
  BrentState st;
  brent_begin(st);
  int k = 0;
  bool more = true;
  for (;;) {
    more = brent_feed_fast(st, -s_ws.spec_ll[k], tol);
    k++;
    if (!more || k >= n || st.u != s_ws.spec_p[k]) break;
  }
  s_ws.brent = st;
  s_ws.n_eval = (unsigned)k;
  s_ws.more = more ? 1 : 0;
  if (more) publish_next(st.u, log2per);
}

template <int U, bool NA, bool ES>
struct WideEval {
  const DevRun *run;
  const uint4 *recs;   // site records in shared memory
  double *spill;       // this block's coefficient scratch for units beyond T*U (5 doubles per unit), or nullptr
  int T, t, cls, n_units;
  double tol;
  int desc[U];         // my units (slot k = unit t + k*T), loaded once per kernel

  // PolymorphismLogLikelihood (FLSeq:91-104) for alleles (a1, a2).  On return (after a block barrier) s_ws.brent holds
  // min / fmin, s_ws.n_eval the rounds, s_ws.h0 the H0 term.
  // with_h0: also the product of the p^4 coefficients (the hom-ref hypothesis under --denovo, main:455-462).
  __device__ __forceinline__ void optimize(int a1, int a2, bool denovo, bool with_h0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (T + 31) >> 5;
    const int g11 = geno_index(a1, a1), g12 = geno_index(a1, a2), g22 = geno_index(a2, a2);
    const int mode = (denovo ? 1 : 0) | (NA ? (cls << 1) : 0);
    // partial products cover PER units each and carry q^(4 PER) (PER = 4 for U = 16: |log2| <= 230 per partial, the
    // four of them multiply to a normal number)
    constexpr int PER = U >= 4 ? U / 4 : 1;
    constexpr int NACC = U / PER;
    constexpr int LOG2PER = PER == 4 ? 2 : (PER == 2 ? 1 : 0);
    // prod over my slots of P_k(r) = 2^(K + Kn) * prod P'_k(r): K from the real units, Kn from the neutral slots
    int K = 0, Kn = 0;
    unsigned fragile = 0;    // bit k: unit k is evaluated the reference's way
    auto account = [&](int code, int k) {
      if (code & (kSlotNeutral | kSlotFragile)) { Kn += 2; if (code & kSlotFragile) fragile |= 1u << k; }
      else K += (code << 12) >> 12;  // sign-extend the 20-bit exponent
    };
    // The out-of-line slot builders write into a staging array in local memory (their results would be spilled around
    // the calls anyway); the coefficients move to registers once, after the last call.
    double stage[U][5];
    if constexpr (NA) {
#pragma unroll
      for (int k = 0; k < U; k++) account(slot_na_ol(recs, desc[k], g11, g12, g22, mode, stage[k]), k);
    } else if constexpr (U == 1) {
      account(slot_single_ol(recs, desc[0], g11, g12, g22, mode, stage[0]), 0);
    } else {
#pragma unroll
      for (int k = 0; k < U; k += 2) {
        // (by value: the ten coefficients come back in registers — measured on B200: 8.7 -> 9.3 M sites/s on 1,000 trios --denovo)
        const PairOut o = slot_pair_val_ol(recs, desc[k], desc[k + 1], g11, g12, g22, mode);
#pragma unroll
        for (int a = 0; a < 5; a++) { stage[k][a] = o.b[0][a]; stage[k + 1][a] = o.b[1][a]; }
        account(o.c0, k); account(o.c1, k + 1);
      }
    }
    double B[U][5];
#pragma unroll
    for (int k = 0; k < U; k++)
#pragma unroll
      for (int a = 0; a < 5; a++) B[k][a] = stage[k][a];
    unsigned spill_fragile = 0;  // more than T*U units: the rest lives in global memory (L2), at most 32 per thread
    if (spill) {
      int j = 0;
      for (int u = T * U + t; u < n_units; u += T, j++) {
        const DevUnit du = run->units[u];
        double tmp[5];
        const int code = NA ? slot_na_ol(recs, desc_pack(du), g11, g12, g22, mode, tmp) : slot_single_ol(recs, desc_pack(du), g11, g12, g22, mode, tmp);
        if (code & kSlotFragile) { spill_fragile |= 1u << j; Kn += 2; } else K += (code << 12) >> 12;
        double *dst = spill + (size_t)(u - T * U) * 5;
        dst[0] = tmp[0]; dst[1] = tmp[1]; dst[2] = tmp[2]; dst[3] = tmp[3]; dst[4] = tmp[4];
      }
    }
    const bool rare = ES || spill != nullptr || (fragile | spill_fragile) != 0;
    RareCtx ctx;
    if (rare) {
      ctx.run = run; ctx.recs = recs; ctx.spill = spill; ctx.fragile = fragile; ctx.spill_fragile = spill_fragile;
      ctx.U = U; ctx.T = T; ctx.t = t; ctx.n_units = n_units; ctx.g11 = g11; ctx.g12 = g12; ctx.g22 = g22; ctx.mode = mode;
      ctx.cls = NA ? cls : PM_CHR_AUTO; ctx.denovo = denovo ? 1 : 0;
#pragma unroll
      for (int k = 0; k < U; k++) ctx.desc[k] = desc[k];
    }
    const int Kall = K + Kn;
    PM_TICK(2);
    // ---- the speculative round: the objective at ALL points of the monotone path (spec_table_ol) in one go — the
    // same evaluations Brent would ask for one by one on a monomorphic site, without a block-wide reduction, a log10
    // and a Brent step between any two of them.  Five points at a time: five independent Horner / shuffle chains.
    constexpr int SPER = U > 8 ? 8 : U;   // units per speculative partial product (which carries q^(4 SPER))
    constexpr int SACC = U / SPER;
    constexpr int SG = 5;
    const int ns = s_ws.spec_n;
    for (int k0 = 0; k0 < ns; k0 += SG) {
      double acc[SG][SACC], rr[SG];
      int kk[SG];
#pragma unroll
      for (int j = 0; j < SG; j++) {
        kk[j] = k0 + j < ns ? k0 + j : ns - 1;  // a short last group repeats the last point (same value, same slot)
        rr[j] = s_ws.spec_r[kk[j]];
        const double qf = s_ws.spec_q[kk[j]];
#pragma unroll
        for (int a = 0; a < SACC; a++) acc[j][a] = qf;
      }
#pragma unroll
      for (int k = 0; k < U; k++)
#pragma unroll
        for (int j = 0; j < SG; j++) acc[j][k / SPER] *= fma(fma(fma(fma(B[k][4], rr[j], B[k][3]), rr[j], B[k][2]), rr[j], B[k][1]), rr[j], B[k][0]);
      double m[SG];
      int e[SG];
      bool zero[SG];
#pragma unroll
      for (int j = 0; j < SG; j++) {
        const double vv = SACC == 2 ? acc[j][0] * acc[j][1] : acc[j][0];
        ME x = me_split_pos(vv);
        x.e += Kall;
        bool z = !(vv > 0.0);
        if (rare) {  // fragile units, units in the L2 scratch, extended families: one out-of-line call per point
          const ME f = rare_factor_ol<ES>(&ctx, s_ws.spec_p[kk[j]], rr[j], 0);
          const double mm = z ? 0.0 : x.m * f.m;
          z = !(mm > 0.0);
          const ME y = me_split_pos(mm);
          x.m = z ? 0.0 : y.m; x.e += f.e + y.e;
        }
        m[j] = x.m; e[j] = x.e; zero[j] = z;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < SG; j++) m[j] *= __shfl_xor_sync(0xffffffffu, m[j], o);
#pragma unroll
      for (int j = 0; j < SG; j++) {
        const int es = __reduce_add_sync(0xffffffffu, e[j]);
        const bool anyz = __any_sync(0xffffffffu, zero[j]);
        if (lane == 0) {
          const ME w = me_split_pos(m[j]);
          s_ws.spec_m[kk[j]][warp] = anyz ? 0.0 : w.m; s_ws.spec_e[kk[j]][warp] = w.e + es;
        }
      }
    }
    if (with_h0) {  // H0 = prod_u B4_u (the likelihood at p = 1), same scaling
      ME h;
      h.m = 1.0; h.e = K;
#pragma unroll
      for (int k = 0; k < U; k++)
        if (desc[k] >= 0 && !(fragile >> k & 1)) me_mul(h, B[k][4]);
      if (rare) {
        const ME f = rare_factor_ol<ES>(&ctx, 1.0, 0.0, 1);
        const double mm = h.m * f.m;
        const ME y = me_split(mm);
        h.m = y.m; h.e += f.e + y.e;
      }
      const bool hz = !(h.m > 0.0);
      double hm = hz ? 1.0 : h.m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hm *= __shfl_xor_sync(0xffffffffu, hm, o);
      const int he = __reduce_add_sync(0xffffffffu, h.e);
      const bool anyz = __any_sync(0xffffffffu, hz);
      if (lane == 0) { const ME w = me_split_pos(hm); s_ws.red0_m[warp] = anyz ? 0.0 : w.m; s_ws.red0_e[warp] = w.e + he; }
    }
    block_sync();
    PM_TICK(6);
    if (warp == 0) {  // lane k: log10 of the block product at point k; lane 31: the H0 product
      if (lane < ns || (with_h0 && lane == 31)) {
        const bool h0 = lane == 31 && lane >= ns;
        double m0 = 1.0, m1 = 1.0;
        int es = 0;
        for (int w = 0; w < nwarp; w++) {
          const double x = h0 ? s_ws.red0_m[w] : s_ws.spec_m[lane][w];
          es += h0 ? s_ws.red0_e[w] : s_ws.spec_e[lane][w];
          if (w & 1) m1 *= x; else m0 *= x;
        }
        const double mm = m0 * m1;
        const ME a = me_split_pos(mm);
        const double ll = mm > 0.0 ? log10_me(a.m, a.e + es) : -CUDART_INF;
        if (h0) s_ws.h0 = ll; else s_ws.spec_ll[lane] = ll;
      }
      __syncwarp();
      if (lane == 0) spec_resolve_ol(tol, LOG2PER);
    }
    block_sync();
    PM_TICK(3);
    if (!s_ws.more) return;
    // Brent left the monotone path (a polymorphic site, or a hypothesis whose first allele is not the common one): one
    // evaluation per round from here.  Two copies of the round loop.  Warp 0's has the call to the out-of-line Brent
    // step: the registers that call clobbers are re-loaded from the stack after it, every round — by warp 0 only.  In
    // the other warps' copy there is no call, and the coefficients never leave their registers.
    if (warp == 0) rounds<true>(B, Kall, rare, ctx);
    else rounds<false>(B, Kall, rare, ctx);
  }

  // ---- H1..H3 (and H0 under --denovo) of an autosomal site in ONE pass ------------------------------------
  // The three (ref, alt) hypotheses share the walk over the records (unit_h123) and the points of the monotone path do
  // not depend on the data, so: unit by unit, the coefficients of the three hypotheses are built in registers and
  // multiplied into 3 x 10 running products straight away — they are never stored — then one block barrier, warp 0 takes
  // the 30 logarithms (lane = hypothesis * 10 + point; lane 31: H0) and checks the three paths with one ballot.  If every
  // hypothesis stays on its path (any monomorphic site) warp 0 calls done(ll1, ll2, ll3, ll_h0) and the site has cost two
  // barriers; otherwise (a path left, a fragile unit, a path longer than ten points) nothing has been written and the
  // caller goes through optimize() hypothesis by hypothesis.  Same evaluation order as optimize(): the same bits.
  // fm / fe: per-warp partials, entry (lane, warp) at lane * nwp + warp.
  static constexpr int kF3Pts = 10;
  template <typename Done>
  __device__ __forceinline__ bool fast3(int ref, bool dn, double *fm, int *fe, int nwp, Done done) {
    constexpr int NP = kF3Pts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (T + 31) >> 5;
    const int ns = s_ws.spec_n;
    if (!s_ws.spec_complete || ns > NP) return false;
    // my units are t, t + T, ...: the first `mine` slots (no neutral slots here: the products start from q^(4 mine))
    const int mine = t < n_units ? min(U, (n_units - t + T - 1) / T) : 0;
    double acc[3][NP];
#pragma unroll
    for (int j = 0; j < NP; j++) {
      const double qf = s_ws.spec_qn[mine][j];  // (slots >= ns repeat the last point)
#pragma unroll
      for (int h = 0; h < 3; h++) acc[h][j] = qf;
    }
    int Kall[3] = {0, 0, 0};
    ME h0;
    h0.m = 1.0; h0.e = 0;
    int bail = 0;
#pragma unroll 1
    for (int k = 0; k < mine; k++) {
      const int4 du = __ldg(reinterpret_cast<const int4 *>(run->units + (t + k * T)));  // first, nkids, kid0, sex
      double c[3][5];
      {
        double b[3][5];
        unit_h123(recs, du.x, du.y, ref, dn ? 1 : 0, b);
#pragma unroll
        for (int h = 0; h < 3; h++) {
          const int code = scale_slot_nb(b[h], du.y >= 0, c[h]);
          bail |= code & kSlotFragile;
          Kall[h] += (code << 12) >> 12;  // sign-extend the 20-bit exponent
        }
      }
      if (dn) me_mul(h0, c[0][4]);
#pragma unroll
      for (int j = 0; j < NP; j++) {
        const double r = s_ws.spec_r[j];
#pragma unroll
        for (int h = 0; h < 3; h++) acc[h][j] *= fma(fma(fma(fma(c[h][4], r, c[h][3]), r, c[h][2]), r, c[h][1]), r, c[h][0]);
      }
    }
    const int K0 = Kall[0];
#pragma unroll
    for (int h = 0; h < 3; h++) {
#pragma unroll
      for (int j0 = 0; j0 < NP; j0 += 5) {
        double m[5];
        int e[5];
        bool zero[5];
#pragma unroll
        for (int j = 0; j < 5; j++) {
          const double vv = acc[h][j0 + j];
          const ME x = me_split_pos(vv);
          zero[j] = !(vv > 0.0);
          m[j] = x.m; e[j] = x.e + Kall[h];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < 5; j++) m[j] *= __shfl_xor_sync(0xffffffffu, m[j], o);
#pragma unroll
        for (int j = 0; j < 5; j++) {
          const int es = __reduce_add_sync(0xffffffffu, e[j]);
          const bool anyz = __any_sync(0xffffffffu, zero[j]);
          if (lane == 0) {
            const ME w = me_split_pos(m[j]);
            fm[(h * NP + j0 + j) * nwp + warp] = anyz ? 0.0 : w.m; fe[(h * NP + j0 + j) * nwp + warp] = w.e + es;
          }
        }
      }
    }
    if (dn) {  // H0 = prod_u B4_u (the likelihood at p = 1), same scaling as H1
      h0.e += K0;
      const bool hz = !(h0.m > 0.0);
      double hm = hz ? 1.0 : h0.m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hm *= __shfl_xor_sync(0xffffffffu, hm, o);
      const int he = __reduce_add_sync(0xffffffffu, h0.e);
      const bool anyz = __any_sync(0xffffffffu, hz);
      if (lane == 0) { const ME w = me_split_pos(hm); fm[31 * nwp + warp] = anyz ? 0.0 : w.m; fe[31 * nwp + warp] = w.e + he; }
    }
    PM_TICK(2);
    if (__syncthreads_or(bail)) return false;
    PM_TICK(6);
    if (warp == 0) {
      const int hh = lane / NP, j = lane - hh * NP;
      double ll = 0.0;
      if (lane < 3 * NP || (dn && lane == 31)) {
        double m0 = 1.0, m1 = 1.0;
        int es = 0;
        for (int w = 0; w < nwarp; w++) {
          const double x = fm[lane * nwp + w];
          es += fe[lane * nwp + w];
          if (w & 1) m1 *= x; else m0 *= x;
        }
        const double mm = m0 * m1;
        const ME a = me_split_pos(mm);
        ll = mm > 0.0 ? log10_me(a.m, a.e + es) : -CUDART_INF;
      }
      const double first = __shfl_sync(0xffffffffu, ll, hh < 3 ? hh * NP : 0);
      const bool ok = lane >= 3 * NP || j == 0 || j >= ns || ll < first;  // f_k > f_0 (a NaN or -inf leaves the path)
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      const double l1 = __shfl_sync(0xffffffffu, ll, 0), l2 = __shfl_sync(0xffffffffu, ll, NP), l3 = __shfl_sync(0xffffffffu, ll, 2 * NP);
      const double lh = __shfl_sync(0xffffffffu, ll, 31);
      const bool good = okmask == 0xffffffffu;
      if (lane == 0) s_ws.ibcast[0] = good ? 1 : 0;
      if (good) done(l1, l2, l3, lh);  // all lanes of warp 0
    }
    block_sync();
    PM_TICK(3);
    return s_ws.ibcast[0] != 0;
  }

  template <bool DRIVER>
  __device__ __forceinline__ void rounds(const double (&B)[U][5], const int Kall, const bool rare, const RareCtx &ctx) {
    constexpr int PER = U >= 4 ? U / 4 : 1;
    constexpr int NACC = U / PER;
    constexpr int LOG2PER = PER == 4 ? 2 : (PER == 2 ? 1 : 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
      const double r = s_ws.next[1], qa = s_ws.next[2];
      double v[NACC];
#pragma unroll
      for (int i = 0; i < NACC; i++) v[i] = qa;
#pragma unroll
      for (int k = 0; k < U; k++) v[k % NACC] *= fma(fma(fma(fma(B[k][4], r, B[k][3]), r, B[k][2]), r, B[k][1]), r, B[k][0]);
      double vv = v[0];
      if (NACC == 4) vv = (v[0] * v[1]) * (v[2] * v[3]);
      else if (NACC == 2) vv = v[0] * v[1];
      ME acc = me_split_pos(vv);
      acc.e += Kall;
      bool zero = !(vv > 0.0);
      if (rare) {
        const ME f = rare_factor_ol<ES>(&ctx, s_ws.next[0], r, 0);
        const double mm = zero ? 0.0 : acc.m * f.m;
        zero = !(mm > 0.0);
        const ME y = me_split_pos(mm);
        acc.m = zero ? 0.0 : y.m; acc.e += f.e + y.e;
      }
      // block product: mantissas by shuffles (each in [1,2): 32 of them cannot overflow), exponents by one REDUX; a
      // zero anywhere makes the whole product zero
      {
        double m = acc.m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m *= __shfl_xor_sync(0xffffffffu, m, o);
        const int e = __reduce_add_sync(0xffffffffu, acc.e);
        const bool anyz = __any_sync(0xffffffffu, zero);
        if (lane == 0) {
          const ME w = me_split_pos(m);
          s_ws.red_m[warp] = anyz ? 0.0 : w.m; s_ws.red_e[warp] = w.e + e;
        }
      }
      block_sync();
      if (DRIVER) {
        PM_TICK(7);
        if (threadIdx.x == 0) brent_round_ol((T + 31) >> 5, tol, LOG2PER);
      }
      block_sync();
      if (DRIVER) PM_TICK(4);
      if (!s_ws.more) break;
    }
  }
};

// NA = false: the autosomal instance; sites on chrX / chrY / MT are left untouched and flagged in err[1].
// NA = true: launched right behind it, returns at once unless err[1] is set, then does only those sites.
// ES = the pedigree also has extended families (evaluated by es_factor); again separate instances.
// GM = the site's records do not fit in shared memory (more than ~13,800 people): they are read where they lie, in
// global memory (L2; the next site of the block is prefetched there as always), no staging copy.
template <int U, int MAXT, int MINB, bool NA, bool ES, bool GM>
__global__ void __launch_bounds__(MAXT, MINB) k_sites_wide(const DevRun *__restrict__ run, const pm_site_hdr *__restrict__ hdr,
                                                           const uint4 *__restrict__ recs_all, const double *__restrict__ mono_all,
                                                           size_t n_sites, double *__restrict__ spill_all, int n_spill, int f3_off,
                                                           pm_site_result *__restrict__ res, uint16_t *__restrict__ status,
                                                           int *__restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WideShared *const ws = &s_ws;
  constexpr int F3W = MAXT <= 32 ? 1 : MAXT / 32 + 1;  // warps per block (+1: an odd stride for the 32 lanes that read the partials); plan_wide knows it too
  if (NA ? ((err[1] == 0 && run->site_filter != 2) || run->site_filter == 1) : run->site_filter == 2) return;
  const int np = run->n_person;
  unsigned char *site_base = smem_raw;  // the dynamic part is the site buffer alone
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = run->lut[i];
  for (int i = threadIdx.x; i < 100; i += blockDim.x) s_mut[i] = run->mut[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { s_log_inv[i] = run->log_inv[i]; s_log_tab[i] = run->log_tab[i]; }
  if (threadIdx.x < 32) { ws->red_m[threadIdx.x] = 1.0; ws->red_e[threadIdx.x] = 0; ws->red0_m[threadIdx.x] = 1.0; ws->red0_e[threadIdx.x] = 0; }
  for (int i = threadIdx.x; i < kMaxSpec * 17; i += blockDim.x) { (&ws->spec_m[0][0])[i] = 1.0; (&ws->spec_e[0][0])[i] = 0; }
  if (threadIdx.x == 0) spec_table_ol(run->precision, U >= 8 ? 3 : (U == 4 ? 2 : (U == 2 ? 1 : 0)));  // q^(4 min(U, 8))
  if (threadIdx.x == 0) {
    for (int k = 0; k < 8; k++) ws->phase[k] = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ws->mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  const uint32_t site_bytes = (uint32_t)np * 16u;
  if (!GM && threadIdx.x == 0 && blockIdx.x < n_sites) tma_issue_site(site_base, recs_all + (size_t)blockIdx.x * np, site_bytes, &ws->mbar);
  WideEval<U, NA, ES> ev;
  ev.run = run; ev.T = blockDim.x; ev.t = threadIdx.x; ev.cls = PM_CHR_AUTO;
  ev.n_units = run->n_units; ev.tol = run->precision;
  ev.recs = reinterpret_cast<const uint4 *>(site_base);
  ev.spill = n_spill > 0 ? spill_all + (size_t)blockIdx.x * n_spill * 5 : nullptr;
#pragma unroll
  for (int k = 0; k < U; k++) {  // my units, packed into one register each for the whole kernel
    const int u = (int)threadIdx.x + k * (int)blockDim.x;
    ev.desc[k] = u < ev.n_units ? desc_pack(run->units[u]) : -1;
  }
  const uint4 *site = reinterpret_cast<const uint4 *>(site_base);

  for (size_t s = blockIdx.x; s < n_sites; s += gridDim.x) {
    const size_t nxt = s + gridDim.x;
    if constexpr (GM) { site = recs_all + s * (size_t)np; ev.recs = site; }
    if (threadIdx.x == 0 && nxt < n_sites) tma_prefetch_l2(recs_all + nxt * (size_t)np, site_bytes);  // next site -> L2 meanwhile
#ifdef PM_PHASE_TIMING
    if (threadIdx.x == 0) ws->t_last = clock64();
#endif
    if constexpr (!GM) mbar_wait(&ws->mbar, phase);
    phase ^= 1;
    PM_TICK(0);
    const pm_site_hdr h = hdr[s];
    const int ref = h.ref_base;
    bool skip = false;
    const int cls = h.chr_class;
    const bool bad_cls = cls > PM_CHR_MT;
    if (NA) ev.cls = cls;
    const bool bad = ref < 1 || ref > 4 || bad_cls;
    unsigned n_hyp = 0, n_eval = 0;
    bool synced = false;  // a block barrier since the wait: every iteration needs one (nobody may run two sites ahead:
                          // the mbarrier's parity tells adjacent phases apart only; the site buffer is refilled behind it)
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
    const bool vcf = run->vcf_mode != 0;
    if (!skip && !vcf) {
      // ---- CalcReadStats / MonomorphismLogLikelihood: integer block reductions, every thread reads the totals ----
      const int grr = geno_index(ref, ref);
      int dsum = 0, nsamp = 0, mq = 0, lksum = 0;
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const uint4 rec = site[i];
        const int d = rec_depth(rec);
        dsum += d; nsamp += d > 0; mq += rec_mapq(rec); lksum += (int)rec_lk(rec, grr);
      }
      dsum = __reduce_add_sync(0xffffffffu, dsum); nsamp = __reduce_add_sync(0xffffffffu, nsamp);
      mq = __reduce_add_sync(0xffffffffu, mq); lksum = __reduce_add_sync(0xffffffffu, lksum);
      const int par = (int)(phase & 1);  // (flipped after the wait: the buffer of this site)
      if (lane == 0) ws->red_i[par][warp] = make_int4(dsum, nsamp, mq, lksum);
      block_sync();
      synced = true;
      // the totals, two divisions and the filters: every thread for itself (no second barrier, no broadcast)
      int D = 0, NS = 0, MQ = 0, LK = 0;
      for (int w = 0; w < nwarp; w++) { const int4 x = ws->red_i[par][w]; D += x.x; NS += x.y; MQ += x.z; LK += x.w; }
      double perc_samp = 0.0, avg_mq = 0.0;
      if (NS > 0) { avg_mq = (double)MQ / (double)NS; perc_samp = (double)NS / (double)np; }
      int st = 0;
      if (D < run->min_total_depth) st = PM_SITE_MIN_DEPTH;
      else if (run->max_total_depth > 0 && D > run->max_total_depth) st = PM_SITE_MAX_DEPTH;
      else if (perc_samp * 100 < run->min_ps) st = PM_SITE_MIN_PS;
      else if (avg_mq < run->min_map_quality) st = PM_SITE_MIN_MAPQ;
      if (threadIdx.x == 0) {
        ws->lk_mono = -(double)LK / 10.0;  // sum_i -lk_i/10 with the integer sum taken first (exact), one division
        pm_site_result &r = ws->r;
        memset(&r, 0, sizeof r);
        r.site = (uint32_t)s; r.maxidx = -1;
        r.total_depth = D; r.num_samp = NS; r.avg_map_qual = avg_mq; r.perc_samp = perc_samp;
        r.status = (uint8_t)st;
        r.reserved = (uint16_t)ref;
        if (st != 0) { res[s] = r; status[s] = status_word(r); }
      }
      PM_TICK(1);
      skip = st != 0;
    }
    if (!skip) {
      // One call site for every minimisation of the site.  step 0: the (REF, ALT) chain of a VCF record (mono is given);
      // 1..6: H1..H6 (H4..H6 only if the posterior after H3 is not decisive, main:499); 7: the mutation-free refit of a
      // polymorphic de novo call (main:567-573).
      const bool dn = run->denovo != 0;
      int step = vcf ? 0 : 1;
      // after H3 / H6 (thread 0): posterior over the hypotheses so far; H4..H6 wanted?  else the decisions of main:539-574
      auto decide = [&](int at) {
        bool more = false;
        if (at == 3) {
          if (!dn) ws->r.varllk[0] = run->cls_log[cls][0] + ws->lk_mono;
          ws->r.varllk_noprior[0] = ws->r.varllk[0] - run->cls_log[cls][0];
          ws->r.varfreq[0] = 1.0;
          var_posterior_ol(&ws->r, ref, 4);
          more = ws->r.var_post_prob < 0.99;  // main:499
        } else {
          var_posterior_ol(&ws->r, ref, 7);
        }
        ws->ibcast[1] = more;
        ws->ibcast[2] = more ? 0 : site_decide_ol(run, &ws->r, ws->lk_mono);
      };
      bool decided = false;
      if constexpr (!NA && !ES) {
        // H1..H3 (+ H0) in one pass wherever every hypothesis stays on Brent's monotone path: any monomorphic site
        // (the per-warp partials of that pass sit behind the site buffer where the block's shared memory has room for them)
        double *const f3_m = reinterpret_cast<double *>(smem_raw + f3_off);
        int *const f3_e = reinterpret_cast<int *>(f3_m + 32 * F3W);
        if (f3_off != 0 && !vcf && ev.spill == nullptr) {
          decided = ev.fast3(ref, dn, f3_m, f3_e, F3W, [&](double l1, double l2, double l3, double lh0) {
            if (lane == 0) {
              PM_TICK(4);
              const double p0 = ws->spec_p[0];
              site_store_hyp(run, ws->r, 1, l1, p0, cls);
              site_store_hyp(run, ws->r, 2, l2, p0, cls);
              site_store_hyp(run, ws->r, 3, l3, p0, cls);
              ws->r.varllk[0] = run->cls_log[cls][0] + (dn ? lh0 : ws->lk_mono);
              ws->r.varllk_noprior[0] = ws->r.varllk[0] - run->cls_log[cls][0];
              ws->r.varfreq[0] = 1.0;
              n_hyp += dn ? 4 : 3; n_eval += 3 * (unsigned)ws->spec_n + (dn ? 1 : 0);
              PM_TICK(7);
            }
            __syncwarp();
            var_posterior_warp_ol(&ws->r, ref, 4);  // (as decide(3), the posterior by the whole warp)
            if (lane == 0) {
              const bool more = ws->r.var_post_prob < 0.99;  // main:499
              ws->ibcast[1] = more;
              ws->ibcast[2] = more ? 0 : site_decide_ol(run, &ws->r, ws->lk_mono);
            }
          });
          if (decided) step = 3;
        }
      }
      if (vcf && threadIdx.x == 0) { memset(&ws->r, 0, sizeof ws->r); ws->r.site = (uint32_t)s; }
      for (;;) {
        if (!decided) {
          int a1, a2;
          bool dnc = false, with_h0 = false;
          if (step == 0) { a1 = ref; a2 = h.reserved & 0xff; }
          else if (step <= 6) { hyp_alleles(step, ref, a1, a2); dnc = dn; with_h0 = dn && step == 1; }
          else { a1 = ws->r.allele1; a2 = ws->r.allele2; }
          ev.optimize(a1, a2, dnc, with_h0);
          // (optimize ends with a block barrier: thread 0's state is final)
          if (threadIdx.x == 0) { n_hyp += with_h0 ? 2 : 1; n_eval += ws->n_eval + (with_h0 ? 1 : 0); }
          if (step == 0) {
            if (threadIdx.x == 0) vcf_record_result(run, ws->r, a1, a2, (h.reserved & 0x100) != 0, mono_all[s], -ws->brent.fmin, ws->brent.min);
            break;
          }
          if (step == 7) {
            if (threadIdx.x == 0) site_finish_refit(run, ws->r, -ws->brent.fmin, ws->brent.min);
            break;
          }
          if (threadIdx.x == 0) {
            site_store_hyp(run, ws->r, step, -ws->brent.fmin, ws->brent.min, cls);
            if (with_h0) ws->r.varllk[0] = run->cls_log[cls][0] + ws->h0;
          }
          if (step != 3 && step != 6) { step++; continue; }
          if (threadIdx.x == 0) decide(step);
          __syncthreads();
        }
        decided = false;
        if (ws->ibcast[1]) step = 4;
        else if (ws->ibcast[2]) step = 7;
        else break;
      }
      // (every way out of the loop above ends with a block barrier behind the last read of the site buffer: the next site's
      // copy starts before thread 0 books the result, the other warps are already waiting for it)
      if (!GM && threadIdx.x == 0 && nxt < n_sites) tma_issue_site(site_base, recs_all + nxt * (size_t)np, site_bytes, &ws->mbar);
      if (threadIdx.x == 0) {
        pm_site_result &r = ws->r;
        if (!vcf) {
          if (r.status == PM_SITE_EMITTED && run->denovo && r.denovo_lr < run->denovo_min_llr) { r.flags |= PM_FLAG_ROW_DROPPED; r.status = PM_SITE_DENOVO_DROPPED; }
          r.reserved = 0;
        }
        status[s] = status_word(r);
        atomicAdd(&run->counters[0], (unsigned long long)n_hyp);
        atomicAdd(&run->counters[1], (unsigned long long)n_eval);
        atomicAdd(&run->counters[2], 1ull);
        atomicAdd(&run->counters[3], (unsigned long long)(r.status == PM_SITE_EMITTED));
#ifdef PM_PHASE_TIMING
        PM_TICK(5);
        for (int k = 0; k < 8; k++) { atomicAdd(&run->counters[8 + k], ws->phase[k]); ws->phase[k] = 0; }
#endif
      }
    }
    // Every path that read the site buffer has passed a block barrier since (the statistics' one, or the one an
    // optimisation ends with): the buffer can be refilled.  ws->r is thread 0's; the other lanes of warp 0 only help to
    // store it.  No block barrier here: the other warps are already waiting for the next site.
    if (skip && !synced) __syncthreads();
    if (warp == 0) {
      __syncwarp();
      if (!GM && skip && lane == 0 && nxt < n_sites) tma_issue_site(site_base, recs_all + nxt * (size_t)np, site_bytes, &ws->mbar);
      if (!skip) {  // the 256-byte result leaves as one coalesced 32 x 8-byte store
        static_assert(sizeof(pm_site_result) == 256, "pm_site_result must be 256 bytes");
        reinterpret_cast<unsigned long long *>(&res[s])[lane] = reinterpret_cast<const unsigned long long *>(&ws->r)[lane];
      }
      __syncwarp();
    }
  }
}

// ================================================================================================
// plans and launchers
// ================================================================================================
static size_t wide_smem_bytes(int n_person) {  // dynamic part = the site buffer; tables and block state are static shared memory (~7 KB)
  return ((((size_t)n_person * 16 + 127) / 128) * 128) + 16;
}

// The (U, MAXT, MINB) instantiations: U = units per thread kept in registers; MAXT = largest block the instance is
// launched with; MINB = resident blocks per SM the register allocation leaves room for.
//   variant 0: U=1 T=32        1..32 units    one warp per site, no cross-warp traffic at all
//   variant 1: U=2 T=32        ..64
//   variant 2: U=4 T=32        ..128
//   variant 3: U=8 T=32        ..256
//   variant 4: U=8 T=64..128   ..1024         168 registers, 3 blocks/SM at T=128 (measured on B200, 1,000 trios --denovo:
//                                             9.7 M sites/s; U=16 T=64: 7.4 M; U=8 T=128 at 128 registers, 4 blocks/SM: 8.3 M)
//   variant 5: U=8 T=160..512  ..4096 (+ the L2 scratch beyond)   128 registers
//   (also measured with the one-pass H1..H3 evaluation: 160 threads x 3 blocks at 96 registers 10.0 M, 256 x 2 at 128 registers
//    8.8 M, 192 x 3 at 112 registers 8.0 M against 13.1 M for 128 x 3 at 168 registers: the spills cost more than the warps bring)
//   variant 6: as 5, the site's records read from global memory (pedigrees whose site does not fit in shared memory)
#define PM_WIDE_VARIANTS(X) X(0, 1, 32, 16, false) X(1, 2, 32, 16, false) X(2, 4, 32, 12, false) X(3, 8, 32, 10, false) X(4, 8, 128, 3, false) X(5, 8, 512, 1, false) X(6, 8, 512, 1, true)
#define X(V_, U_, MT_, MB_, GM_) U_,
static const int kVariantU[] = {PM_WIDE_VARIANTS(X)};
#undef X
#define X(V_, U_, MT_, MB_, GM_) MT_,
static const int kVariantMaxT[] = {PM_WIDE_VARIANTS(X)};
#undef X
static const int kNumVariants = (int)(sizeof(kVariantU) / sizeof(kVariantU[0]));

// The dynamic shared-memory limit of a kernel is a property of the (instantiation, device) pair, shared by every context
// of the process: it is raised to everything the SM offers beside the kernel's static part, never to one pedigree's
// need (a second engine with a smaller pedigree on the same instantiation must not lower it under the first one).
template <typename K>
static cudaError_t wide_raise_smem(K kernel, size_t smem) {
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa;
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  const long long room = (long long)optin - (long long)fa.sharedSizeBytes;
  if ((long long)smem > room) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room);
}
template <int U, int MAXT, int MINB, bool GM>
static cudaError_t wide_attr(bool es, size_t smem, int threads, int *per_sm) {
  cudaError_t e;
  if (es) {
    e = wide_raise_smem(k_sites_wide<U, MAXT, MINB, true, true, GM>, smem);
    if (e == cudaSuccess) e = wide_raise_smem(k_sites_wide<U, MAXT, MINB, false, true, GM>, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_sites_wide<U, MAXT, MINB, false, true, GM>, threads, smem);
  } else {
    e = wide_raise_smem(k_sites_wide<U, MAXT, MINB, true, false, GM>, smem);
    if (e == cudaSuccess) e = wide_raise_smem(k_sites_wide<U, MAXT, MINB, false, false, GM>, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_sites_wide<U, MAXT, MINB, false, false, GM>, threads, smem);
  }
  return e;
}

template <int U, int MAXT, int MINB, bool GM>
static void wide_launch(const LaunchPlan &plan, unsigned grid, size_t smem, cudaStream_t stream, const DevRun *d_run, const pm_site_hdr *d_hdr,
                        const uint4 *d_recs, const double *d_mono, size_t n_sites, double *d_spill, pm_site_result *d_res, uint16_t *d_status,
                        int *d_err) {
  if (plan.es) {
    k_sites_wide<U, MAXT, MINB, false, true, GM><<<grid, plan.threads, smem, stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, plan.n_spill, plan.f3_offset, d_res, d_status, d_err);
    k_sites_wide<U, MAXT, MINB, true, true, GM><<<grid, plan.threads, smem, stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, plan.n_spill, plan.f3_offset, d_res, d_status, d_err);
  } else {
    k_sites_wide<U, MAXT, MINB, false, false, GM><<<grid, plan.threads, smem, stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, plan.n_spill, plan.f3_offset, d_res, d_status, d_err);
    k_sites_wide<U, MAXT, MINB, true, false, GM><<<grid, plan.threads, smem, stream>>>(d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, plan.n_spill, plan.f3_offset, d_res, d_status, d_err);
  }
}

cudaError_t launch_sites_wide(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs, const double *d_mono,
                              size_t n_sites, double *d_spill, pm_site_result *d_res, uint16_t *d_status, int *d_err, cudaStream_t stream) {
  const size_t smem = (size_t)plan.smem_bytes;
  const unsigned grid = (unsigned)(n_sites < (size_t)plan.grid ? n_sites : (size_t)plan.grid);
  cudaError_t e = cudaMemsetAsync(d_err + 1, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  switch (plan.variant) {
#define X(V_, U_, MT_, MB_, GM_) case V_: wide_launch<U_, MT_, MB_, GM_>(plan, grid, smem, stream, d_run, d_hdr, d_recs, d_mono, n_sites, d_spill, d_res, d_status, d_err); break;
    PM_WIDE_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// Chooses threads per block T and the instantiation for a pedigree with n_units quartic units.
cudaError_t plan_wide(LaunchPlan *plan, int n_person, int n_units, int n_es, int sm_count, const int *force /* {variant, threads} or nullptr */) {
  plan->kind = LaunchPlan::WIDE;
  plan->n_person = n_person;
  plan->es = n_es > 0 ? 1 : 0;  // extended families ride along as thread-serial peels (es_factor)
  auto up32 = [](int x) { return ((x + 31) / 32) * 32; };
  int variant, T, U;
  if (n_units <= 32) { variant = 0; U = 1; T = 32; }
  else if (n_units <= 64) { variant = 1; U = 2; T = 32; }
  else if (n_units <= 128) { variant = 2; U = 4; T = 32; }
  else if (n_units <= 256) { variant = 3; U = 8; T = 32; }
  else if (n_units <= 1024) { variant = 4; U = 8; T = up32((n_units + 7) / 8); }
  else { variant = 5; U = 8; T = up32((n_units + 7) / 8); if (T > 512) T = 512; }
  if (force) {
    if (force[0] < 0 || force[0] >= kNumVariants || force[1] < 32 || force[1] % 32 || force[1] > kVariantMaxT[force[0]]) return cudaErrorInvalidValue;
    variant = force[0]; T = force[1]; U = kVariantU[variant];
  }
  plan->variant = variant;
  plan->threads = T;
  plan->units_per_thread = U;
  plan->n_spill = n_units > T * U ? n_units - T * U : 0;
  if (plan->n_spill > 32 * T) return cudaErrorNotSupported;  // the fragile-unit mask of the spilled units is 32 bits per thread
  // Shared memory per block: the site buffer and, behind it, the 32 x (warps + 1) partial products (8 + 4 bytes) of the
  // one-pass H1..H3 evaluation.  A pedigree whose site nearly fills the SM's shared memory goes without that pass (same
  // results, one hypothesis at a time); one whose site does not fit at all (more than ~13,800 people) takes the
  // instantiation that reads the records from global memory.
  int per_sm = 1;
  cudaError_t e = cudaErrorInvalidValue;
  auto fit = [&](int v, size_t site_smem) {
    const int f3w = kVariantMaxT[v] <= 32 ? 1 : kVariantMaxT[v] / 32 + 1;
    const size_t f3_bytes = (size_t)32 * f3w * 12;
    for (int with_f3 = 1; with_f3 >= 0; with_f3--) {
      const size_t smem = site_smem + (with_f3 ? f3_bytes : 0);
      per_sm = 0;
      switch (v) {
#define X(V_, U_, MT_, MB_, GM_) case V_: e = wide_attr<U_, MT_, MB_, GM_>(plan->es != 0, smem, T, &per_sm); break;
        PM_WIDE_VARIANTS(X)
#undef X
      }
      plan->smem_bytes = (int)smem;
      plan->f3_offset = with_f3 ? (int)site_smem : 0;
      if (e == cudaSuccess && per_sm >= 1) return true;
      (void)cudaGetLastError();
    }
    return false;
  };
  constexpr int kGlobalVariant = 6, kLargestVariant = 5;
  bool ok = variant == kGlobalVariant ? fit(variant, 16) : fit(variant, wide_smem_bytes(n_person));
  if (!ok && variant == kLargestVariant) {
    variant = kGlobalVariant;
    plan->variant = variant;
    ok = fit(variant, 16);
  }
  if (!ok) return e != cudaSuccess ? e : cudaErrorNotSupported;
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
#ifdef PM_PHASE_TIMING  // measurement builds only (scripts/gpu_phase_timing.py): fewer resident blocks, to tell latency from contention
  if (const char *env = getenv("PM_BLOCKS_PER_SM")) { const int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
#endif
  plan->grid = sm_count * per_sm;  // persistent: a multiple of the SM count
  plan->blocks_per_sm = per_sm;
  return cudaSuccess;
}

}  // namespace pm
