/*
 * pm_oracle.h — CPU restatement of polymutt's per-site family-likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (polymutt_b200/, the C-ABI library, the CLI)
 * may include, link or execute this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  The restatement (plus the host front end and VCF writers) reproduces the
 * four golden VCFs shipped in the reference's example/ directory byte for byte (non-## lines), and
 * reproduces byte for byte the outputs of the unmodified reference built by oracle/build_ref.sh on
 * extended pedigrees, --denovo, --all_sites, --pos, --quick_call, chrX / chrY / MT, multi-section GLFs and
 * VCF input (tests/test_oracle_golden.py, fixtures + generating script under tests/golden/).
 *
 * It shares the data contract (structs) of include/polymutt_b200.h so that results can be
 * compared field by field with the CUDA path.
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H

#include "../include/polymutt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmo_ctx pmo_ctx;

pmo_ctx *pmo_create(const pm_pedigree *ped, const pm_params *par, const double *lut256);
void pmo_destroy(pmo_ctx *ctx);

/* Mirrors pm_call_glf_sites with out_mode = PM_OUT_ALL: res_out[n_sites],
 * person_out[n_sites*n_person] (rows of non-emitted sites are zeroed), status_out[n_sites]. */
int pmo_call_glf_sites(pmo_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *person_site,
                       size_t n_sites, uint16_t *status_out, pm_site_result *res_out,
                       pm_person_result *person_out);

/* Mirrors pm_call_vcf_records (ctx created with pm_params.vcf_input = 1 and the PL2LK table as lut256). */
int pmo_call_vcf_records(pmo_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono,
                         size_t n_records, pm_site_result *res_out, pm_person_result *person_out);

/* ES_Peeling restated (src/FamilyLikelihoodES.cpp:46-277). Same contract as pm_build_peel_order. */
int pmo_build_peel_order(int32_t n, const int32_t *father, const int32_t *mother, const uint8_t *sex,
                         pm_peel_step *steps);

/* One family likelihood at a fixed frequency (unit-test hook):
 * log10 L_f(freq) for alleles (a1,a2) on the site loaded last by pmo_load_site. */
int pmo_load_site(pmo_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *persons);
double pmo_family_loglik(pmo_ctx *ctx, int fam, int a1, int a2, double freq, int denovo);
double pmo_all_family_loglik(pmo_ctx *ctx, int a1, int a2, double freq, int denovo);
/* Brent as OptimizeFrequency runs it (src/NucFamGenotypeLikelihood.cpp:432-444); returns maxlogL, *freq = min */
double pmo_optimize(pmo_ctx *ctx, int a1, int a2, int denovo, double *freq, int *n_eval);

void pmo_fill_lut(double *lut256);
void pmo_genotype_mutation_matrix(double mu, double tstv, double *m100);
const char *pmo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
