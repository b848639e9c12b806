/* Shim: bgzf is only used by tabix RANGE_MODE and the bgzip VCF writer, neither of which polymutt
 * reaches. Opening a bgzf stream fails. Ours, not reference code. */
#ifndef PM_SHIM_BGZF_H
#define PM_SHIM_BGZF_H
typedef struct { int dummy; } BGZF;
static inline BGZF *bgzf_open(const char *fn, const char *mode) { (void)fn; (void)mode; return 0; }
static inline int bgzf_close(BGZF *fp) { (void)fp; return 0; }
static inline int bgzf_write(BGZF *fp, const void *data, int length) { (void)fp; (void)data; (void)length; return -1; }
#endif
