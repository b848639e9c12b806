// The per-chromosome control loop of the drop-in executable (the role of src/main.cpp:248-627):
// load the pedigree once, merge the per-person GLF streams into packed site batches, hand each batch
// to a likelihood engine, print VCF rows in site order and the per-chromosome summary block.
//
// The engine is reached through a small table of C function pointers.  The product executable binds
// it to the CUDA C-ABI (pm_create / pm_call_glf_sites / ...); nothing else is linked into it.
#pragma once
#include <cstddef>
#include <cstdint>

#include "polymutt_b200.h"

namespace pmh {

struct Engine {
  const char *name;
  void *(*create)(const pm_pedigree *, const pm_params *, const double *lut, int device);
  // Same contract as pm_call_glf_sites with out_mode = PM_OUT_EMITTED.
  int (*call_glf)(void *ctx, const pm_site_hdr *, const pm_person_site *, size_t n_sites, uint16_t *status,
                  pm_site_result *res, pm_person_result *person, size_t res_cap, size_t *n_res);
  void (*destroy)(void *ctx);
  const char *(*last_error)();
  // Same contract as pm_call_vcf_records (--in_vcf); nullptr if the engine has no VCF-input path.
  int (*call_vcf)(void *ctx, const pm_site_hdr *, const pm_person_site *, const double *mono, size_t n, pm_site_result *res,
                  pm_person_result *person) = nullptr;
  // optional: same with the per-sample results reduced to calls[n * n_person] = best | gq << 8 (pm_call_vcf_records_calls):
  // all the --in_vcf writer prints from; used when present
  int (*call_vcf_calls)(void *ctx, const pm_site_hdr *, const pm_person_site *, const double *mono, size_t n, pm_site_result *res,
                        uint16_t *calls) = nullptr;
  // optional: page-locked allocation for the batch buffers (nullptr = plain malloc)
  void *(*host_alloc)(size_t bytes) = nullptr;
  void (*host_free)(void *p) = nullptr;
  // optional: pm_call_vcf_records_pl -- three PL bytes per sample in (a1a1, a1a2, a2a2), calls out; used when present
  int (*call_vcf_pl)(void *ctx, const pm_site_hdr *, const uint8_t *pl3, const double *mono, size_t n, pm_site_result *res,
                     uint16_t *calls) = nullptr;
};

// Returns the process exit code (0 on success, 1 after a fatal error, like the reference's error()).
int run_cli(int argc, char **argv, const Engine &engine);

}  // namespace pmh
