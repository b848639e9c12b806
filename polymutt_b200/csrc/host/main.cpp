// polymutt-b200: drop-in executable for polymutt's GLF-input calling path.  Same command line,
// .ped/.dat/GLF-index/GLF inputs and VCF output as the reference (src/main.cpp); the per-site
// likelihood engine is the CUDA C-ABI library — there is no CPU engine in this binary.
#include "driver.h"

static void *create(const pm_pedigree *ped, const pm_params *par, const double *lut, int device) {
  return pm_create(ped, par, lut, device);
}
static int call_glf(void *ctx, const pm_site_hdr *hdr, const pm_person_site *ps, size_t n, uint16_t *status,
                    pm_site_result *res, pm_person_result *person, size_t cap, size_t *n_res) {
  return pm_call_glf_sites((pm_ctx *)ctx, hdr, ps, n, PM_OUT_EMITTED, status, res, person, cap, n_res);
}
static int call_vcf(void *ctx, const pm_site_hdr *hdr, const pm_person_site *ps, const double *mono, size_t n, pm_site_result *res,
                    pm_person_result *person) {
  return pm_call_vcf_records((pm_ctx *)ctx, hdr, ps, mono, n, res, person);
}
static int call_vcf_calls(void *ctx, const pm_site_hdr *hdr, const pm_person_site *ps, const double *mono, size_t n, pm_site_result *res,
                          uint16_t *calls) {
  return pm_call_vcf_records_calls((pm_ctx *)ctx, hdr, ps, mono, n, res, calls);
}
static int call_vcf_pl(void *ctx, const pm_site_hdr *hdr, const uint8_t *pl3, const double *mono, size_t n, pm_site_result *res, uint16_t *calls) {
  return pm_call_vcf_records_pl((pm_ctx *)ctx, hdr, pl3, mono, n, res, calls);
}
static void destroy(void *ctx) { pm_destroy((pm_ctx *)ctx); }

int main(int argc, char **argv) {
  pmh::Engine e{"cuda-sm100a", create, call_glf, destroy, pm_last_error, call_vcf, call_vcf_calls, pm_host_alloc, pm_host_free, call_vcf_pl};
  return pmh::run_cli(argc, argv, e);
}
