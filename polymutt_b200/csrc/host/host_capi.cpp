// Host-only entry points of the C ABI (no GPU needed): pm_last_error, pm_fill_lut,
// pm_genotype_mutation_matrix.  pm_build_peel_order lives in peel_order.cpp.
#include <cmath>

#include "host_error.h"
#include "polymutt_b200.h"

extern "C" const char *pm_last_error(void) { return pmh::last_error(); }
extern "C" int pm_abi_version(void) { return PM_ABI_VERSION; }

// 256-entry table of 10^(-i/10), computed with the expression the reference uses so that the
// doubles are bit-identical to glfHandler's likelihoods (core/BaseQualityHelper.cpp:12-13).
extern "C" void pm_fill_lut(double *lut256) {
  for (int i = 0; i <= 255; i++) lut256[i] = pow(0.1, i * 0.1);
}

// Genotype mutation matrix of the de novo model (src/MutationModel.cpp:15-30, 46-90): per-allele
// 4x4 matrix with rate mu split by the ts/tv ratio, its Kronecker square over ordered genotypes,
// then the two orderings of every heterozygote merged on the "to" side.
extern "C" void pm_genotype_mutation_matrix(double mu, double tstv, double *m100) {
  double a[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) a[i][j] = (i == j) ? 1 - mu : (1 - mu) / 3;
  if (tstv != 0.0) {
    const double ts = mu / 3 * (3 - 3 / (1 + tstv)), tv = mu / 3 * (0.5 / (1 + tstv) * 3);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        if (i != j) a[i][j] = ((i ^ j) == 2) ? ts : tv;  // A<->G, C<->T are transitions
  }
  // unordered genotype g <-> ordered pairs (x<=y) in AA AC AG AT CC CG CT GG GT TT order
  int gx[10], gy[10], g = 0;
  for (int x = 0; x < 4; x++) for (int y = x; y < 4; y++) { gx[g] = x; gy[g] = y; g++; }
  for (int from = 0; from < 10; from++)
    for (int to = 0; to < 10; to++) {
      int i = gx[from], j = gy[from], x = gx[to], y = gy[to];
      double v = a[i][x] * a[j][y];
      if (x != y) v += a[i][y] * a[j][x];
      m100[from * 10 + to] = v;
    }
}
