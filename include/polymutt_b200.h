/*
 * polymutt_b200.h — C ABI of the B200-native per-site family-likelihood engine.
 *
 * The reference (polymutt 0.13) has no plugin/FFI seam: its drivers (src/main.cpp:325-594 for GLF
 * input, src/PedVCF.cpp:116-163 for VCF input) call public methods of the C++ class
 * FamilyLikelihoodSeq{,_VCF} once per site and read its public members
 * (src/NucFamGenotypeLikelihood.h:24-76).  This header replaces exactly that class surface with a
 * batched, stream-ordered C interface: the host front end (ours) does what
 * PedigreeGLF::Move2NextBaseEntry + FillPenetrance do (src/PedigreeGLF.cpp:282-324,
 * src/FamilyLikelihoodSeq.cpp:296-317) and packs many sites; the library does everything from
 * CalcReadStats (src/NucFamGenotypeLikelihood.cpp:520) through CalcPostProb (src/main.cpp:576-587)
 * on the GPU; OutputVCF* stays on the host and consumes pm_site_result / pm_person_result.
 *
 * Plain C, no C++ types, no exceptions.  Every call returns PM_OK (0) or a negative PM_E* code and
 * leaves a message retrievable by pm_last_error().  The caller owns every host buffer; the library
 * owns all device memory.  One pm_ctx is bound to one CUDA device and one stream; use one ctx (and
 * one host thread) per GPU.  There is no CPU fallback: pm_create fails if no sm_100 device is usable.
 */
#ifndef POLYMUTT_B200_H
#define POLYMUTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_ABI_VERSION 1

/* status codes */
#define PM_OK            0
#define PM_EINVAL       -1   /* bad argument / unsupported pedigree */
#define PM_ECUDA        -2   /* CUDA runtime error (message has the cudaError string) */
#define PM_ENOMEM       -3
#define PM_EUNSUPPORTED -4   /* feature the device path does not implement yet (fails loudly) */

/* genotype order everywhere: AA AC AG AT CC CG CT GG GT TT (core/glfHandler.h:102-106);
 * bases 1..4 = A C G T (core/glfHandler.cpp:4 translateBase). */
#define PM_NGENO 10
#define PM_MAX_HYP 7

/* chromosome class of a site: selects the site prior and the transmission rules
 * (src/main.cpp:312-315, NucFamGenotypeLikelihood::SetPolyPrior* :231-293). */
enum { PM_CHR_AUTO = 0, PM_CHR_X = 1, PM_CHR_Y = 2, PM_CHR_MT = 3 };

/* peeling step kinds (src/FamilyLikelihoodES.h:36, "peelingType") */
enum { PM_PEEL_CHILD_TO_PARENTS = 1, PM_PEEL_SPOUSE_TO_SPOUSE = 2, PM_PEEL_PARENTS_TO_CHILD = 3 };

/* One step of the Elston–Stewart order built by ES_Peeling (src/FamilyLikelihoodES.cpp:135-277).
 * Person indices are positions inside the family (Family::path order, founders first).
 *   type 1: from0 = child,            to0,to1 = (father, mother)
 *   type 2: from0 = spouse peeled,    to0 = spouse kept
 *   type 3: from0,from1 = the couple as the reference stores it, to0 = child
 * Unused slots are -1. */
typedef struct pm_peel_step {
  int32_t type;
  int32_t from0, from1;
  int32_t to0, to1;
} pm_peel_step;

/* Pedigree topology, in VCF column order: families in sorted famid order, members in
 * Family::path order (founders first; src/NucFamGenotypeLikelihood.cpp:1777-1784).
 * All arrays are caller-owned and copied by pm_create. */
typedef struct pm_pedigree {
  int32_t n_fam;
  int32_t n_person;              /* = sum(fam_size) */
  const int32_t *fam_size;       /* [n_fam]  Family::count */
  const int32_t *fam_founders;   /* [n_fam]  Family::founders */
  const int32_t *fam_generations;/* [n_fam]  Family::generations (1,2,3); nuclear <=> generations==2 && founders==2 */
  const uint8_t *sex;            /* [n_person] 0 unknown, 1 male, 2 female */
  const int32_t *father;         /* [n_person] in-family index of the father, -1 for founders */
  const int32_t *mother;         /* [n_person] in-family index of the mother, -1 for founders */
  const int32_t *peel_first;     /* [n_fam+1] offsets into peel[]; empty range for founder-only families */
  const pm_peel_step *peel;      /* concatenated peeling orders; may be NULL when every family is
                                    nuclear or founders-only */
} pm_pedigree;

/* Calling parameters = the CmdLinePar fields the likelihood engine reads (src/CmdLinePar.h,
 * defaults from src/main.cpp:59-85). */
typedef struct pm_params {
  double theta;              /* --theta            1e-3 */
  double theta_indel;        /* --indel_theta      1e-4 (VCF mode only) */
  double poly_tstv;          /* --poly_tstv        2.0  */
  double posterior_cutoff;   /* -c                 0.5  */
  double precision;          /* --prec             1e-4 (Brent relative tolerance) */
  double denovo_mut_rate;    /* --rate_denovo      1.5e-8 */
  double denovo_tstv;        /* --tstv_denovo      2.0 */
  double denovo_min_llr;     /* --minLLR_denovo    0.01 (raw ratio, not log10) */
  double min_ps;             /* --minPercSampleWithData (percent) */
  int32_t min_map_quality;   /* --minMapQuality */
  int32_t min_total_depth;   /* --minDepth */
  int32_t max_total_depth;   /* --maxDepth (0 = off) */
  int32_t denovo;            /* --denovo */
  int32_t force_call;        /* set by --pos */
  int32_t out_all_sites;     /* --all_sites */
  int32_t quick_call;        /* --quick_call */
  int32_t vcf_input;         /* 1 = this ctx serves pm_call_vcf_records (--in_vcf, src/PedVCF.cpp) */
} pm_params;

/* Per-site header, 8 bytes. */
typedef struct pm_site_hdr {
  uint32_t pos;        /* 0-based position (PedigreeGLF::currentPos); the VCF prints pos+1 */
  uint8_t  ref_base;   /* 0 = N/other (site is skipped, src/main.cpp:340), 1..4 = A C G T */
  uint8_t  chr_class;  /* PM_CHR_* */
  uint16_t reserved;   /* GLF input: PM_HDR_* bits; VCF input: ALT allele | indel << 8 */
} pm_site_hdr;

/* GLF input on chrX / chrY / MT without --denovo: the reference's nuclear-family code reads a stale member (`sex`,
 * src/NucFamGenotypeLikelihood.h:72) that the genotype-posterior loops of the PREVIOUS emitted site left behind
 * (NucFam.cpp:600-610 vs 1211-1261).  From the second emitted site of a run on that is a constant of the pedigree
 * (the sex of the last person of the last family), which is what the engine assumes.  For the one site whose
 * posteriors are the first the process computes, set this bit and the engine uses the initial value 0 instead. */
#define PM_HDR_FIRST_POSTPROB 0x1

/* Per-(site, person) record, 16 bytes, site-major / person-interleaved:
 * record (s, i) lives at person_site + 16*(s*n_person + i).
 * A person with no GLF file, or whose stream has no record at this position, is all zeros
 * (likelihood 1.0 for every genotype, depth 0: core/glfHandler.cpp:279-317). */
typedef struct pm_person_site {
  uint8_t lk[PM_NGENO];   /* glfEntry::lk, -10*log10 likelihood ratio, genotype order above */
  uint8_t depth[3];       /* glfEntry::depth, 24-bit little endian */
  uint8_t map_quality;    /* glfEntry::mapQuality */
  uint8_t pad[2];
} pm_person_site;

/* The same record without the two pad bytes, 14 bytes: what a host buffer needs to carry per (site, person).
 * record (s, i) lives at wire + 14*(s*n_person + i).  pm_call_glf_sites_wire copies this form over PCIe (12.5 % fewer
 * bytes on a link-bound path) and widens it to pm_person_site on the device (k_unpack_wire). */
typedef struct pm_person_site_wire {
  uint8_t lk[PM_NGENO];   /* glfEntry::lk */
  uint8_t depth[3];       /* glfEntry::depth, 24-bit little endian */
  uint8_t map_quality;    /* glfEntry::mapQuality */
} pm_person_site_wire;

/* What happened to a site (the `continue`s of src/main.cpp:339-574, in order). */
enum {
  PM_SITE_EMITTED       = 0,  /* a VCF row is due */
  PM_SITE_BAD_REF       = 1,  /* refBase not in 1..4                       main.cpp:340 */
  PM_SITE_MIN_DEPTH     = 2,  /* totalDepth < --minDepth                   main.cpp:345 */
  PM_SITE_MAX_DEPTH     = 3,  /* totalDepth > --maxDepth                   main.cpp:346 */
  PM_SITE_MIN_PS        = 4,  /* percSampWithData*100 < --minPerc...       main.cpp:347 */
  PM_SITE_MIN_MAPQ      = 5,  /* avgMapQual < --minMapQuality              main.cpp:348 */
  PM_SITE_NOCALL        = 6,  /* varPostProb < posterior cutoff            main.cpp:539 */
  PM_SITE_MONO          = 7,  /* best hypothesis is hom-ref, nothing to print  main.cpp:555 */
  PM_SITE_DENOVO_LOW_LR = 8,  /* --denovo, mono winner, DQ <= log10(minLLR)    main.cpp:563 */
  PM_SITE_QUICK_SKIP    = 9,  /* --quick_call pre-pass said no variant     main.cpp:432-433 */
  PM_SITE_DENOVO_DROPPED = 10 /* --denovo: the site reached OutputVCF_denovo, which prints nothing because
                                 denovoLR < the raw --minLLR_denovo (src/NucFamGenotypeLikelihood.cpp:1868;
                                 main.cpp:563 compared against log10 of it).  The reference computes genotype
                                 posteriors for these sites and throws them away; we do not compute them.
                                 Only visible effect: the VCF header is printed once such a site is seen. */
};

#define PM_FLAG_NOCALL       0x1  /* varPostProb < cutoff was counted ("Hard to call"), main.cpp:539 */
#define PM_FLAG_ROW_DROPPED  0x2  /* set together with PM_SITE_DENOVO_DROPPED */
#define PM_FLAG_MONO         0x4  /* emitted as monomorphic (isMono / denovo_mono) */

/* Per-site result = the public members of famlk[0] the VCF writers and the summary block read
 * (src/NucFamGenotypeLikelihood.h:24-76, src/main.cpp:596-614). 256 bytes. */
typedef struct pm_site_result {
  uint32_t site;             /* index of the site inside the call's batch */
  uint8_t  status;           /* PM_SITE_* */
  int8_t   maxidx;           /* argmax hypothesis 0..6 (CalcVarPosterior), -1 if not evaluated */
  uint8_t  allele1, allele2; /* famlk[0].allele1/2 at output time (1..4) */
  uint8_t  n_hyp;            /* 0, 4 or 7 hypotheses evaluated */
  uint8_t  flags;            /* PM_FLAG_* */
  uint16_t reserved;
  int32_t  total_depth;      /* totalDepth */
  int32_t  num_samp;         /* numSampWithData */
  double   perc_samp;        /* percSampWithData (fraction) */
  double   avg_map_qual;     /* avgMapQual */
  double   var_post_prob;    /* varPostProb */
  double   poly_qual;        /* polyQual; the VCF prints int(polyQual+0.5) */
  double   freq;             /* GetMinimizer() = famlk[0].min at output time (AF) */
  double   denovo_lr;        /* denovoLR (log10), DQ */
  double   ab;               /* AB */
  double   varllk[PM_MAX_HYP];         /* varllk[0..6] */
  double   varllk_noprior[PM_MAX_HYP]; /* varllk_noprior[0..6] */
  double   varfreq[PM_MAX_HYP];        /* varfreq[0..6] */
  double   refit_llk;        /* --denovo, polymorphic winner: maxlogL refitted without mutation (main.cpp:570) */
} pm_site_result;

/* Per-(emitted site, person) result: postProb[f][j][0..9], bestGenoIdx, GQ, dosage. 96 bytes.
 * post[0..2] are (11, 12, 22) posteriors unless ten_state != 0, where post[0..9] follow the
 * ten-genotype order (kids and extended pedigrees under --denovo). */
typedef struct pm_person_result {
  double  post[PM_NGENO];
  double  dosage;
  int32_t best;       /* bestGenoIdx */
  uint8_t gq;         /* int(-10*log10(1-post[best])+0.5) capped at 100 (NucFam.cpp:1819-1820) */
  uint8_t ten_state;
  uint8_t reserved[2];
} pm_person_result;

typedef struct pm_ctx pm_ctx;

/* lut256[i] must be pow(0.1, i*0.1) computed on the host exactly as
 * core/BaseQualityHelper.cpp:12-13 does; pass NULL to let the library compute it the same way. */
pm_ctx *pm_create(const pm_pedigree *ped, const pm_params *par, const double *lut256, int device);
void    pm_destroy(pm_ctx *ctx);

/* Selects which sites get a pm_site_result / pm_person_result back. */
#define PM_OUT_EMITTED 0   /* only PM_SITE_EMITTED sites, compacted in site order (production) */
#define PM_OUT_ALL     1   /* every site, index == site (parity tests); per-person rows only valid for emitted sites */

/* GLF-input calling for a batch of sites with HOST buffers (includes H2D/D2H copies).
 *   hdr          [n_sites]
 *   person_site  [n_sites * n_person] 16-byte records
 *   status_out   [n_sites] PM_SITE_* | (maxidx+1)<<4 | PM_FLAG_NOCALL<<8 ; may be NULL
 *   res_out      capacity res_cap records; receives *n_res records
 *   person_out   capacity res_cap * n_person records (row r belongs to res_out[r]); may be NULL
 * Returns PM_EINVAL if res_cap is too small (then *n_res holds the required count). */
int pm_call_glf_sites(pm_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *person_site,
                      size_t n_sites, int out_mode, uint16_t *status_out,
                      pm_site_result *res_out, pm_person_result *person_out, size_t res_cap,
                      size_t *n_res);

/* The same call on 14-byte records (pm_person_site_wire): same results, 14/16 of the bytes over PCIe.  Replaces the
 * same reference loop (src/main.cpp:325-594 over PedigreeGLF::glf[i].data, core/glfHandler.h:21-42). */
int pm_call_glf_sites_wire(pm_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site_wire *person_site_wire,
                           size_t n_sites, int out_mode, uint16_t *status_out,
                           pm_site_result *res_out, pm_person_result *person_out, size_t res_cap,
                           size_t *n_res);

/* Same computation on buffers that already live in this ctx's device memory (cudaMalloc pointers,
 * e.g. torch tensors' data_ptr()).  Asynchronous on the ctx stream; call pm_sync before reading. */
int pm_call_glf_sites_device(pm_ctx *ctx, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site,
                             size_t n_sites, int out_mode, uint16_t *d_status_out,
                             pm_site_result *d_res_out, pm_person_result *d_person_out,
                             size_t res_cap, uint32_t *d_n_res);

int pm_sync(pm_ctx *ctx);

/* VCF-input calling (src/PedVCF.cpp:116-163, src/FamilyLikelihoodSeq_VCF.cpp) for a batch of bi-allelic
 * records, HOST buffers.  The ctx must have been created with pm_params.vcf_input = 1 and, as lut256, the
 * table 10^(-i/10) computed as pow(10, -double(i)/10.0) (FamilyLikelihoodSeq_VCF.cpp:21-22).
 *   hdr[r]      pos = POS, ref_base = allele1 (1..4; 1 for an indel), reserved = allele2 | (indel << 8)
 *   person_site records in VCF-column-of-the-pedigree order; lk[g(a1,a1)], lk[g(a1,a2)], lk[g(a2,a2)] hold
 *               int(PL) (or int(-10*GL)) capped at 255, every other byte 0 (a sample that is not in the VCF
 *               or has no data is all zeros = likelihood 1, FamilyLikelihoodSeq_VCF.cpp:275-279)
 *   mono[r]     sum over samples of loglk[ref/ref] (MonomorphismLogLikelihood, FamilyLikelihoodSeq_VCF.cpp:74-83),
 *               computed by the caller because it uses the un-truncated PL/GL doubles
 * Every record produces a row: res_out[n], person_out[n * n_person].  res_out[r].poly_qual is QUAL,
 * .freq the frequency of allele1 (the VCF prints AF = 1 - freq), varllk[0] / varllk[1] are llk_ref / llk_alt. */
int pm_call_vcf_records(pm_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono,
                        size_t n_records, pm_site_result *res_out, pm_person_result *person_out);

/* The same with the per-sample results reduced to what the --in_vcf writer prints from (src/FamilyLikelihoodSeq_VCF.cpp:
 * 499-517: GT from bestGenoIdx, GQ): calls_out[n_records * n_person] = best | gq << 8.  2 bytes per sample come back
 * over PCIe instead of 96. */
int pm_call_vcf_records_calls(pm_ctx *ctx, const pm_site_hdr *hdr, const pm_person_site *person_site, const double *mono,
                              size_t n_records, pm_site_result *res_out, uint16_t *calls_out);

/* The same from the wire form of a VCF record's likelihoods: pl3[n_records * n_person * 3] holds, per sample, int(PL) (or
 * int(-10*GL)) capped at 255 of the genotypes (a1,a1), (a1,a2), (a2,a2) -- the three numbers the reference reads from a
 * sample's PL / GL field (src/FamilyLikelihoodSeq_VCF.cpp:340-382); 0,0,0 for a sample without data.  3 bytes per sample
 * go over PCIe instead of 16; the device widens them to pm_person_site records (k_unpack_pl3). */
int pm_call_vcf_records_pl(pm_ctx *ctx, const pm_site_hdr *hdr, const uint8_t *pl3, const double *mono,
                           size_t n_records, pm_site_result *res_out, uint16_t *calls_out);

/* The same on buffers in this ctx's device memory (asynchronous on the ctx stream, pm_sync before reading).  Every
 * record gets a row: d_res_out[n_records], d_person_out[n_records * n_person], d_status_out[n_records].
 * has_nonauto: the batch holds chrX / chrY / MT records (they take a second pass over the batch). */
int pm_call_vcf_records_device(pm_ctx *ctx, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                               size_t n_records, int has_nonauto, uint16_t *d_status_out, pm_site_result *d_res_out,
                               pm_person_result *d_person_out);

/* Device-buffer variant of pm_call_vcf_records_calls: d_calls_out[n_records * n_person] = best | gq << 8. */
int pm_call_vcf_records_calls_device(pm_ctx *ctx, const pm_site_hdr *d_hdr, const pm_person_site *d_person_site, const double *d_mono,
                                     size_t n_records, int has_nonauto, uint16_t *d_status_out, pm_site_result *d_res_out,
                                     uint16_t *d_calls_out);

/* Page-locked host memory for the buffers handed to pm_call_glf_sites (lets its H2D/D2H copies overlap
 * the kernels).  Optional: pageable buffers are accepted too. */
void *pm_host_alloc(size_t bytes);
void  pm_host_free(void *p);

/* Device time (ms, CUDA events on the ctx stream) and launch count of the calling kernels
 * in the most recent pm_call_*; used by bench.py for the roofline numbers. */
int pm_last_timing(pm_ctx *ctx, float *ms_main_kernel, float *ms_total, int *n_launches);

/* Human-readable description of the kernel plan chosen for this pedigree (bench.py prints it). */
int pm_describe_plan(pm_ctx *ctx, char *buf, size_t len);

/* Test / tuning hook, not needed in production: re-plans the main pass on instantiation `variant` of the block-per-site
 * kernel (pm_wide.cu, PM_WIDE_VARIANTS) with `threads` threads per block, also for pedigrees small enough for the
 * thread-per-site kernel.  The parity tests use it to put every edge case through every kernel. */
int pm_force_wide_plan(pm_ctx *ctx, int variant, int threads);

/* Device-side stopwatch on the ctx stream (CUDA events): pm_timer_start records an event, pm_timer_stop
 * records a second one, waits for it and returns the elapsed milliseconds.  bench.py brackets its K
 * timed steps with these so that the number is taken on the stream the kernels are launched on. */
int pm_timer_start(pm_ctx *ctx);
int pm_timer_stop(pm_ctx *ctx, float *ms);

/* Work counters accumulated by the site kernels since the last pm_reset_counters: hypotheses set up
 * (coefficient builds), objective evaluations (Brent steps incl. the bracket point), sites evaluated
 * (passed the filters) — the inputs of the algorithmic flop count in DESIGN.md. */
typedef struct pm_counters {
  unsigned long long hypotheses;
  unsigned long long evaluations;
  unsigned long long sites_evaluated;
  unsigned long long sites_emitted;
} pm_counters;
int pm_get_counters(pm_ctx *ctx, pm_counters *out);
int pm_reset_counters(pm_ctx *ctx);

/* Microbenchmarks used as roofline denominators: sustained DFMA throughput (FLOP/s) and
 * device copy bandwidth (B/s) measured on this ctx's device. */
int pm_measure_fp64_peak(pm_ctx *ctx, double *flops);
int pm_measure_copy_bw(pm_ctx *ctx, double *bytes_per_s);

const char *pm_last_error(void);
int pm_abi_version(void);

/* ---- host helpers (no GPU needed) ------------------------------------------------------- */

/* Builds the Elston–Stewart peeling order of one family exactly as ES_Peeling does
 * (src/FamilyLikelihoodES.cpp:46-277).  father/mother/sex are in-family arrays of length n.
 * steps must have room for n entries; returns the number of steps written or a negative PM_E*. */
int pm_build_peel_order(int32_t n, const int32_t *father, const int32_t *mother, const uint8_t *sex,
                        pm_peel_step *steps);

/* pow(0.1, i*0.1), i = 0..255 (core/BaseQualityHelper.cpp:12-13). */
void pm_fill_lut(double *lut256);

/* The 10x10 genotype mutation matrix (src/MutationModel.cpp:15-90), row = true genotype. */
void pm_genotype_mutation_matrix(double mu, double tstv, double *m100);

#ifdef __cplusplus
}
#endif
#endif /* POLYMUTT_B200_H */
