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
  int units_per_thread;   // wide: units whose coefficients a thread keeps in registers (template parameter U)
  int variant;            // wide: which (U, MAXT, MINB) instantiation (pm_wide.cu)
  int n_spill;            // wide: units beyond threads * units_per_thread, coefficients in a global-memory scratch
  int grid;               // wide: persistent grid (multiple of the SM count); narrow: derived from n_sites
  int blocks_per_sm;
  int es;                 // wide: the pedigree has extended families too (the ES instances of the kernel)
  int n_person;
  int smem_bytes;         // wide: dynamic shared memory per block (site buffer [+ the one-pass partials])
  int f3_offset;          // wide: byte offset of the one-pass H1..H3 partials in it, 0 = no room for them (that path is off)
  int ten_state;          // narrow: the run evaluates ten-state (--denovo) peels; set by the caller after plan_launch (which zeroes the plan)
};

// force_wide (tests, tuning scripts): {variant, threads} overrides the choice and sends even small pedigrees to the wide kernel
cudaError_t plan_launch(LaunchPlan *plan, int n_person, int n_units, int n_es, int sm_count, const int *force_wide);
cudaError_t plan_wide(LaunchPlan *plan, int n_person, int n_units, int n_es, int sm_count, const int *force);

// d_spill: plan.grid * plan.n_spill * 5 doubles when plan.n_spill > 0, else unused
cudaError_t launch_sites(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                         const double *d_mono, size_t n_sites, double *d_spill, pm_site_result *d_res, uint16_t *d_status, int *d_err,
                         cudaStream_t stream);
cudaError_t launch_sites_wide(const LaunchPlan &plan, const DevRun *d_run, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                              const double *d_mono, size_t n_sites, double *d_spill, pm_site_result *d_res, uint16_t *d_status,
                              int *d_err, cudaStream_t stream);

// --quick_call: sites the unrelated pre-pass did not call (no-call or hom-ref there) become PM_SITE_QUICK_SKIP
cudaError_t launch_quick_merge(const uint16_t *d_status_q, size_t n_sites, pm_site_result *d_res, uint16_t *d_status, cudaStream_t stream);

// d_tile_scratch: (n_sites + 1023) / 1024 words (the per-tile counts of the three-launch form used for long batches)
cudaError_t launch_compact(const uint16_t *d_status, size_t n_sites, uint32_t *d_emit_sites, uint32_t *d_n_emit, int all,
                           uint32_t *d_tile_scratch, cudaStream_t stream);

cudaError_t launch_post(const DevRun *d_run, int n_fam, const pm_site_hdr *d_hdr, const uint4 *d_recs,
                        const pm_site_result *d_res_all, const uint32_t *d_emit_sites, const uint32_t *d_n_emit,
                        size_t max_rows, size_t res_cap, pm_site_result *d_res_out, pm_person_result *d_person_out,
                        uint16_t *d_calls_out, bool has_es, bool ten_state, int sm_count, bool with_ab, cudaStream_t stream);  // ten_state: --denovo on GLF input  // d_calls_out != nullptr: 2 bytes per person instead of d_person_out's 96

// 14-byte wire records -> 16-byte records (d_wire: n_recs * 14 bytes, 16-byte aligned, readable up to the next multiple of 16)
cudaError_t launch_unpack_wire(const void *d_wire, uint4 *d_recs, size_t n_recs, int sm_count, cudaStream_t stream);

// VCF input: 3 PL bytes per (record, sample) -> 16-byte records (genotype indices from the record's alleles in d_hdr)
cudaError_t launch_unpack_pl3(const pm_site_hdr *d_hdr, const uint8_t *d_pl3, uint4 *d_recs, size_t n_records, int np, int sm_count, cudaStream_t stream);

cudaError_t launch_dfma_peak(double *d_out, int blocks, int threads, int iters, cudaStream_t stream);
cudaError_t launch_copy(const void *src, void *dst, size_t bytes, int sm_count, cudaStream_t stream);

}  // namespace pm
