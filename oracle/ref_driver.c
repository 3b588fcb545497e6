/* ref_driver.c -- TEST/BENCH INFRASTRUCTURE ONLY.
 *
 * Multi-threaded batch runner for a CPU implementation of the ksw2 hot path that lives in
 * ANOTHER shared object (dlopen'ed by path): either oracle/_ref/libksw2_ref.so (the unmodified
 * reference, symbols ksw_ext{z,d,s}2_sse + kalloc) or oracle/libksw2_oracle.so (the restatement,
 * symbols kso_ext{z,d,s}2).  Used (a) by tests to get all result fields + CIGARs of many pairs in
 * one call and (b) by bench.py for the cpu_baseline / --impl reference arm: static contiguous
 * sharding over `nthreads` pthreads, one kalloc arena per thread when the library has km_init
 * (BASELINE.md section 3), wall clock around the alignment calls only.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <dlfcn.h>
#include <pthread.h>
#include <time.h>
#include "../include/ksw2.h"

typedef void (*fn_z)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int, int, int, int, ksw_extz_t*);
typedef void (*fn_d)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int8_t, int8_t, int, int, int, int, ksw_extz_t*);
typedef void (*fn_s)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int8_t, int8_t, int, int8_t, int, const uint8_t*, ksw_extz_t*);

typedef void (*fn_rz)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int, int, int, ksw_extz_t*);                    /* ksw_extz */
typedef void (*fn_rd)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int8_t, int8_t, int, int, int, ksw_extz_t*);    /* ksw_extd */

typedef void (*fn_f)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, int8_t, int8_t, int, int, ksw_extz_t*);                                          /* ksw_extf2_sse */

typedef int (*fn_g)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int, int*, int*, uint32_t**);           /* ksw_gg / ksw_gg2 / ksw_gg2_sse */

typedef struct {
	int kind;                 /* 6, 7, 8: the global-alignment entry points ksw_gg / ksw_gg2 / ksw_gg2_sse (flag & 1: no CIGAR pointers); 0 extz2, 1 extd2, 2 exts2, 3 ksw_extz (row-wise), 4 ksw_extd (row-wise), 5 extf2 (q = mch, q2 = mis, zdrop = xdrop) */
	int m; const int8_t *mat;
	int q, e, q2, e2;         /* exts2: q2 = gapo2, e2 unused */
	int w, zdrop, end_bonus, flag, noncan, junc_bonus;
} ksd_params_t;

#define KSD_NF 12 /* fields per pair: max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar reach_end m_cigar */

typedef struct {
	void *fn; void *(*km_init)(void); void (*km_destroy)(void*); void (*kfree)(void*, void*);
	const ksd_params_t *P;
	int64_t lo, hi;
	const uint8_t *qcat, *tcat, *jcat; const int64_t *qoff, *toff;
	int32_t *res; uint32_t **cig;  /* cig[i]: malloc'ed copy (or NULL) */
	int64_t *cells; int64_t (*last_cells)(void);
	int repeat;
	pthread_barrier_t *bar;
	const int32_t *wv;        /* band per pair (NULL: P->w for all) */
	int64_t *next, n;         /* wv != NULL: pairs are handed out dynamically, 8 at a time (mixed lengths: static shards would be unbalanced) */
} work_t;

static void *worker(void *arg)
{
	work_t *W = (work_t*)arg;
	const ksd_params_t *P = W->P;
	void *km = W->km_init ? W->km_init() : 0;
	ksw_extz_t ez; int64_t i; int rep;
	memset(&ez, 0, sizeof ez);
	pthread_barrier_wait(W->bar);
	for (rep = 0; rep < W->repeat; ++rep)
	for (;;) {
		int64_t lo = W->lo, hi = W->hi, g;
		if (W->wv) { g = __atomic_fetch_add(W->next, 8, __ATOMIC_RELAXED); if (g >= W->n * (rep + 1)) break; lo = g - W->n * rep; hi = lo + 8 < W->n ? lo + 8 : W->n; if (lo < 0) lo = 0; }
	for (i = lo; i < hi; ++i) {
		const uint8_t *qs = W->qcat + W->qoff[i], *ts = W->tcat + W->toff[i];
		int ql = (int)(W->qoff[i + 1] - W->qoff[i]), tl = (int)(W->toff[i + 1] - W->toff[i]);
		const int w = W->wv ? W->wv[i] : P->w;
		if (P->kind == 0) ((fn_z)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, w, P->zdrop, P->end_bonus, P->flag, &ez);
		else if (P->kind == 1) ((fn_d)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, (int8_t)P->q2, (int8_t)P->e2, w, P->zdrop, P->end_bonus, P->flag, &ez);
		else if (P->kind >= 6) {
			const int sc = ((fn_g)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, P->w,
			                             (P->flag & 1) ? 0 : &ez.m_cigar, (P->flag & 1) ? 0 : &ez.n_cigar, (P->flag & 1) ? 0 : &ez.cigar);
			ez.max = 0; ez.zdropped = 0; ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1; ez.mqe = ez.mte = KSW_NEG_INF; ez.reach_end = 0;
			ez.score = sc; if (P->flag & 1) ez.n_cigar = 0;
		}
		else if (P->kind == 5) ((fn_f)W->fn)(km, ql, qs, tl, ts, (int8_t)P->q, (int8_t)P->q2, (int8_t)P->e, P->w, P->zdrop, &ez);
		else if (P->kind == 3) ((fn_rz)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, P->w, P->zdrop, P->flag, &ez);
		else if (P->kind == 4) ((fn_rd)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, (int8_t)P->q2, (int8_t)P->e2, P->w, P->zdrop, P->flag, &ez);
		else ((fn_s)W->fn)(km, ql, qs, tl, ts, (int8_t)P->m, P->mat, (int8_t)P->q, (int8_t)P->e, (int8_t)P->q2, (int8_t)P->noncan, P->zdrop, (int8_t)P->junc_bonus, P->flag, W->jcat ? W->jcat + W->toff[i] : 0, &ez);
		if (W->res) {
			int32_t *o = W->res + i * KSD_NF;
			o[0] = (int32_t)ez.max; o[1] = ez.zdropped; o[2] = ez.max_q; o[3] = ez.max_t; o[4] = ez.mqe; o[5] = ez.mqe_t;
			o[6] = ez.mte; o[7] = ez.mte_q; o[8] = ez.score; o[9] = ez.n_cigar; o[10] = ez.reach_end; o[11] = ez.m_cigar;
		}
		if (W->cells && W->last_cells) W->cells[i] = W->last_cells();
		if (W->cig && rep == 0) {
			W->cig[i] = 0;
			if (ez.n_cigar > 0) {
				W->cig[i] = (uint32_t*)malloc((size_t)ez.n_cigar * 4);
				memcpy(W->cig[i], ez.cigar, (size_t)ez.n_cigar * 4);
			}
		}
	}
		if (!W->wv) break;
	}
	pthread_barrier_wait(W->bar);
	if (ez.cigar) { if (km && W->kfree) W->kfree(km, ez.cigar); else free(ez.cigar); }
	if (km && W->km_destroy) W->km_destroy(km);
	return 0;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* Runs pairs [0,n) (sequences concatenated, offsets have n+1 entries; jcat uses toff) with `nthreads`
 * threads, `repeat` passes.  res: n*KSD_NF int32 or NULL.  cig_off/cig_buf: if non-NULL, CIGARs are
 * concatenated into cig_buf (capacity cig_cap words) with offsets cig_off[n+1]; returns -needed if too small.
 * *seconds gets the wall time of the alignment calls.  cells (optional, n entries): executed in-band cells per pair when
 * the library reports them (the oracle port does, the reference cannot).  Returns 0 on success, >0 on load errors. */
int64_t ksd_run_w(const char *libpath, const char *symbol, const ksd_params_t *P, int64_t n,
                  const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *jcat, const int32_t *wv,
                  int nthreads, int repeat, int32_t *res, int64_t *cig_off, uint32_t *cig_buf, int64_t cig_cap, double *seconds, int64_t *cells);
int64_t ksd_run(const char *libpath, const char *symbol, const ksd_params_t *P, int64_t n,
                const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *jcat,
                int nthreads, int repeat, int32_t *res, int64_t *cig_off, uint32_t *cig_buf, int64_t cig_cap, double *seconds, int64_t *cells)
{
	return ksd_run_w(libpath, symbol, P, n, qcat, qoff, tcat, toff, jcat, 0, nthreads, repeat, res, cig_off, cig_buf, cig_cap, seconds, cells);
}
/* the same with a band per pair (wv[i] replaces P->w; only meaningful for kinds 0 and 1) and dynamic hand-out of the pairs */
int64_t ksd_run_w(const char *libpath, const char *symbol, const ksd_params_t *P, int64_t n,
                  const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *jcat, const int32_t *wv,
                  int nthreads, int repeat, int32_t *res, int64_t *cig_off, uint32_t *cig_buf, int64_t cig_cap, double *seconds, int64_t *cells)
{
	int64_t next = 0;
	void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
	void *fn; int t; pthread_t *th; work_t *W; pthread_barrier_t bar; uint32_t **cig = 0; double t0, t1; int64_t i, tot = 0;
	if (!h) { fprintf(stderr, "ksd_run: dlopen(%s): %s\n", libpath, dlerror()); return 1; }
	fn = dlsym(h, symbol);
	if (!fn) { fprintf(stderr, "ksd_run: no symbol %s in %s\n", symbol, libpath); return 2; }
	if (nthreads < 1) nthreads = 1;
	if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
	if (repeat < 1 || wv) repeat = 1;   /* (the dynamic hand-out counts through one pass) */
	if (cig_off) cig = (uint32_t**)calloc((size_t)(n > 0 ? n : 1), sizeof(uint32_t*));
	th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads); W = (work_t*)calloc(nthreads, sizeof(work_t));
	pthread_barrier_init(&bar, 0, nthreads + 1);
	for (t = 0; t < nthreads; ++t) {
		W[t].fn = fn; W[t].P = P; W[t].lo = n * t / nthreads; W[t].hi = n * (t + 1) / nthreads;
		W[t].km_init = (void*(*)(void))dlsym(h, "km_init"); W[t].km_destroy = (void(*)(void*))dlsym(h, "km_destroy");
		W[t].kfree = (void(*)(void*, void*))dlsym(h, "kfree");
		W[t].qcat = qcat; W[t].tcat = tcat; W[t].jcat = jcat; W[t].qoff = qoff; W[t].toff = toff;
		W[t].res = res; W[t].cig = cig; W[t].repeat = repeat; W[t].bar = &bar;
		W[t].cells = cells; W[t].last_cells = (int64_t(*)(void))dlsym(h, "kso_last_cells");
		W[t].wv = wv; W[t].next = &next; W[t].n = n;
		pthread_create(&th[t], 0, worker, &W[t]);
	}
	pthread_barrier_wait(&bar); t0 = now_s();
	pthread_barrier_wait(&bar); t1 = now_s();
	for (t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
	pthread_barrier_destroy(&bar);
	if (seconds) *seconds = t1 - t0;
	if (cig_off) {
		for (i = 0; i < n; ++i) { cig_off[i] = tot; tot += res ? res[i * KSD_NF + 9] : 0; }
		cig_off[n] = tot;
		if (tot <= cig_cap && cig_buf)
			for (i = 0; i < n; ++i) if (cig[i]) memcpy(cig_buf + cig_off[i], cig[i], (size_t)(cig_off[i + 1] - cig_off[i]) * 4);
		for (i = 0; i < n; ++i) free(cig[i]);
		free(cig);
	}
	free(th); free(W);
	/* keep the library loaded: dlclose is skipped on purpose (cheap, avoids re-init churn) */
	return (cig_off && tot > cig_cap) ? -tot : 0;
}
