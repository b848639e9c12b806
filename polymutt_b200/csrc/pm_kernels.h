// pm_kernels.h — host-visible launch interface of pm_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "polymutt_b200.h"

namespace pm {

struct DevRun;

constexpr int kNarrowMaxUnits = 8;  // quartic units a single thread keeps in registers

struct LaunchPlan {
  enum Kind { NARROW = 0, WIDE = 1 } kind;
  int threads;            // block size
  int units_per_thread;   // wide: template parameter U
  int chains;             // wide: Brent chains optimised concurrently (template parameter NC)
  int grid;               // wide: persistent grid (multiple of the SM count); narrow: derived from n_sites
  int blocks_per_sm;
  int site_buffers;       // wide: 2 = the next site is prefetched by TMA while this one is computed
  int low_regs;           // wide: use the 128-register instantiation (more resident blocks per SM)
  int kid_table;          // wide, --denovo: kids' ten mutation-mixed likelihoods are built once per site in shared memory
  int es;                 // wide: the pedigree has extended families too (the ES instances of the kernel)
  int n_person;
};

cudaError_t plan_launch(LaunchPlan *plan, int n_person, int n_units, int n_es, int n_kids_denovo, int sm_count);

cudaError_t launch_sites(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                         const double *d_mono, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, int *d_err,
                         cudaStream_t stream);

// --quick_call: sites the unrelated pre-pass did not call (no-call or hom-ref there) become PM_SITE_QUICK_SKIP
cudaError_t launch_quick_merge(const uint16_t *d_status_q, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, cudaStream_t stream);

cudaError_t launch_compact(const uint16_t *d_status, size_t n_sites, uint32_t *d_emit_sites, uint32_t *d_n_emit, int all,
                           cudaStream_t stream);

cudaError_t launch_post(const DevRun *d_run, int n_fam, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                        const pm_site_result *d_res_all, const uint32_t *d_emit_sites, const uint32_t *d_n_emit,
                        size_t max_rows, size_t res_cap, pm_site_result *d_res_out, pm_person_result *d_person_out,
                        int sm_count, cudaStream_t stream);

cudaError_t launch_dfma_peak(double *d_out, int blocks, int threads, int iters, cudaStream_t stream);
cudaError_t launch_copy(const void *src, void *dst, size_t bytes, int sm_count, cudaStream_t stream);

}  // namespace pm
