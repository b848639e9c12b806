/* Shim for the absent bzlib.h: base/IO.h wraps bz2 streams, which the polymutt hot path never opens.
 * Any attempt to use bz2 aborts. Ours, not reference code. */
#ifndef PM_SHIM_BZLIB_H
#define PM_SHIM_BZLIB_H
#include <stdio.h>
#include <stdlib.h>
typedef void BZFILE;
#define BZ_OK 0
#define BZ_STREAM_END 4
static inline BZFILE *BZ2_bzReadOpen(int *e, FILE *f, int v, int s, void *u, int n) { (void)f;(void)v;(void)s;(void)u;(void)n; if (e) *e = -1; fprintf(stderr, "bz2 input is not supported in the oracle build\n"); abort(); return 0; }
static inline void BZ2_bzclose(BZFILE *b) { (void)b; }
static inline void BZ2_bzReadClose(int *e, BZFILE *b) { (void)e; (void)b; }
static inline int BZ2_bzRead(int *e, BZFILE *b, void *buf, int len) { (void)b;(void)buf;(void)len; if (e) *e = -1; return 0; }
static inline BZFILE *BZ2_bzWriteOpen(int *e, FILE *f, int b, int v, int w) { (void)f;(void)b;(void)v;(void)w; if (e) *e = -1; abort(); return 0; }
static inline void BZ2_bzWrite(int *e, BZFILE *b, void *buf, int len) { (void)e;(void)b;(void)buf;(void)len; }
static inline void BZ2_bzWriteClose(int *e, BZFILE *b, int a, unsigned *i, unsigned *o) { (void)e;(void)b;(void)a;(void)i;(void)o; }
#endif
