/* Shim for the absent tabix-0.2.5 (reference third/README.tgi). polymutt only opens VCFs in
 * LINE_MODE (libVcf/VCFInputFile.h:163-243); RANGE_MODE is never reached, so every ti_* call
 * here reports failure. Ours, not reference code. */
#ifndef PM_SHIM_TABIX_H
#define PM_SHIM_TABIX_H
typedef struct { int dummy; } ti_index_t;
typedef struct { ti_index_t *idx; } tabix_t;
typedef struct { int dummy; } *ti_iter_t;
static inline tabix_t *ti_open(const char *fn, const char *idx) { (void)fn; (void)idx; return 0; }
static inline int ti_lazy_index_load(tabix_t *t) { (void)t; return -1; }
static inline void ti_close(tabix_t *t) { (void)t; }
static inline ti_iter_t ti_querys(tabix_t *t, const char *reg) { (void)t; (void)reg; return 0; }
static inline const char *ti_read(tabix_t *t, ti_iter_t it, int *len) { (void)t; (void)it; if (len) *len = 0; return 0; }
static inline void ti_iter_destroy(ti_iter_t it) { (void)it; }
static inline ti_iter_t ti_query(tabix_t *t, const char *name, int beg, int end) { (void)t; (void)name; (void)beg; (void)end; return 0; }
static inline ti_iter_t ti_queryi(tabix_t *t, int tid, int beg, int end) { (void)t; (void)tid; (void)beg; (void)end; return 0; }
static inline int ti_parse_region(const ti_index_t *idx, const char *str, int *tid, int *beg, int *end) { (void)idx; (void)str; (void)tid; (void)beg; (void)end; return -1; }
#endif
