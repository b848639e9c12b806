// TEST INFRASTRUCTURE ONLY.  The host front end of polymutt_b200 (flag parser, pedigree loader, GLF
// merge, VCF writers) driven by the CPU oracle instead of the CUDA library.  Its VCF text is diffed
// against the reference's golden files and against the stock binary built by oracle/build_ref.sh:
// that pins the oracle AND the host front end to the reference.  Never shipped, never linked into the
// product library or executable.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../polymutt_b200/csrc/host/driver.h"
#include "pm_oracle.h"

struct OracleCtx { pmo_ctx *c; int np; };

static void *create(const pm_pedigree *ped, const pm_params *par, const double *lut, int) {
  pmo_ctx *c = pmo_create(ped, par, lut);
  if (!c) return nullptr;
  return new OracleCtx{c, ped->n_person};
}
static int call_glf(void *vctx, const pm_site_hdr *hdr, const pm_person_site *ps, size_t n, uint16_t *status,
                    pm_site_result *res, pm_person_result *person, size_t cap, size_t *n_res) {
  OracleCtx *o = (OracleCtx *)vctx;
  std::vector<pm_site_result> all(n);
  std::vector<pm_person_result> pall(n * (size_t)o->np);
  int rc = pmo_call_glf_sites(o->c, hdr, ps, n, status, all.data(), pall.data());
  if (rc) return rc;
  size_t k = 0;
  for (size_t s = 0; s < n; s++)
    if (all[s].status == PM_SITE_EMITTED) {
      if (k < cap) {
        res[k] = all[s];
        memcpy(&person[k * (size_t)o->np], &pall[s * (size_t)o->np], sizeof(pm_person_result) * (size_t)o->np);
      }
      k++;
    }
  *n_res = k;   // like pm_call_glf_sites: on overflow the needed row count comes back with PM_EINVAL
  return k > cap ? PM_EINVAL : PM_OK;
}
static int call_vcf(void *vctx, const pm_site_hdr *hdr, const pm_person_site *ps, const double *mono, size_t n, pm_site_result *res,
                    pm_person_result *person) {
  return pmo_call_vcf_records(((OracleCtx *)vctx)->c, hdr, ps, mono, n, res, person);
}
static void destroy(void *vctx) { OracleCtx *o = (OracleCtx *)vctx; pmo_destroy(o->c); delete o; }

int main(int argc, char **argv) {
  pmh::Engine e{"cpu-oracle", create, call_glf, destroy, pmo_last_error, call_vcf};
  return pmh::run_cli(argc, argv, e);
}
