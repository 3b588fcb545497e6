/* kalloc_interop.c -- TEST INFRASTRUCTURE ONLY (built into oracle/_ref/ by oracle/Makefile, where /root/reference exists).
 *
 * The ownership rule of the drop-in boundary (reference ksw2.h:103-119, kalloc.c:136): ez->cigar is grown with the CALLER's
 * krealloc(km, ...) on the calling thread and freed by the caller with kfree(km, ez.cigar).  This program is such a caller: it links the
 * reference's kalloc.c and the reference kernels statically, loads libksw2_b200.so with dlopen and, from several threads that each own
 * a private kalloc arena (km_init), aligns the same random pairs with both, compares every field and CIGAR word, re-uses the ez buffers
 * across calls and frees them with kfree; then the same through the array-of-pointers batch call ksw2b_extd2_batch(ctx, km, ...).  The GPU library finds the program's krealloc with dlsym(RTLD_DEFAULT) (-rdynamic).
 * usage: kalloc_interop <path to libksw2_b200.so> [threads] [pairs per thread]      exit 0 = all equal
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksw2.h"
#include "kalloc.h"

typedef void (*fn_d)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int8_t, int8_t, int, int, int, int, ksw_extz_t*);
typedef void (*fn_z)(void*, int, const uint8_t*, int, const uint8_t*, int8_t, const int8_t*, int8_t, int8_t, int, int, int, int, ksw_extz_t*);
typedef void *(*fn_create)(int);
typedef void (*fn_destroy)(void*);
typedef int (*fn_dbatch)(void*, void*, int64_t, const int*, const uint8_t *const*, const int*, const uint8_t *const*, int8_t, const int8_t*, int8_t, int8_t, int8_t, int8_t, int, int, int, int, ksw_extz_t*);
static fn_d gpu_extd2; static fn_z gpu_extz2;
static fn_create gpu_create; static fn_destroy gpu_destroy; static fn_dbatch gpu_extd2_batch;
static int n_pairs = 40;
static int8_t mat[25];

typedef struct { int tid; int bad; long calls; } arg_t;

static uint64_t rnd(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

static void *worker(void *p)
{
	arg_t *a = (arg_t*)p;
	void *km = km_init();                       /* this thread's arena: not thread-safe, never shared */
	uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(a->tid + 1);
	ksw_extz_t er, eg;
	uint8_t *t = (uint8_t*)malloc(4096), *q = (uint8_t*)malloc(4096);
	memset(&er, 0, sizeof er); memset(&eg, 0, sizeof eg);
	for (int i = 0; i < n_pairs; ++i) {
		const int tl = 30 + (int)(rnd(&s) % 1500), dual = (int)(rnd(&s) & 1), flag = (int[]){0, KSW_EZ_RIGHT, KSW_EZ_EXTZ_ONLY, KSW_EZ_REV_CIGAR}[rnd(&s) & 3];
		const int w = (rnd(&s) & 1) ? -1 : 20 + (int)(rnd(&s) % 200), zdrop = (rnd(&s) & 1) ? -1 : 100 + (int)(rnd(&s) % 300);
		int ql = 0;
		for (int k = 0; k < tl; ++k) t[k] = (uint8_t)(rnd(&s) & 3);
		for (int k = 0; k < tl && ql < 4000; ++k) {
			const unsigned u = (unsigned)(rnd(&s) % 100);
			if (u < 3) continue;
			if (u < 6) q[ql++] = (uint8_t)(rnd(&s) & 3);
			q[ql++] = u > 92 ? (uint8_t)((t[k] + 1) & 3) : t[k];
		}
		if (ql == 0) q[ql++] = 0;
		if (dual) {
			ksw_extd2_sse(km, ql, q, tl, t, 5, mat, 4, 2, 24, 1, w, zdrop, 0, flag, &er);
			gpu_extd2(km, ql, q, tl, t, 5, mat, 4, 2, 24, 1, w, zdrop, 0, flag, &eg);
		} else {
			ksw_extz2_sse(km, ql, q, tl, t, 5, mat, 4, 2, w, zdrop, 0, flag, &er);
			gpu_extz2(km, ql, q, tl, t, 5, mat, 4, 2, w, zdrop, 0, flag, &eg);
		}
		++a->calls;
		if (er.max != eg.max || er.zdropped != eg.zdropped || er.max_q != eg.max_q || er.max_t != eg.max_t || er.mqe != eg.mqe || er.mqe_t != eg.mqe_t ||
		    er.mte != eg.mte || er.mte_q != eg.mte_q || er.score != eg.score || er.n_cigar != eg.n_cigar || er.reach_end != eg.reach_end ||
		    (er.n_cigar > 0 && memcmp(er.cigar, eg.cigar, (size_t)er.n_cigar * 4) != 0)) {
			fprintf(stderr, "thread %d pair %d: mismatch (dual %d flag %d w %d zdrop %d tl %d ql %d): score %d vs %d, n_cigar %d vs %d\n",
			        a->tid, i, dual, flag, w, zdrop, tl, ql, er.score, eg.score, er.n_cigar, eg.n_cigar);
			++a->bad;
		}
		if (eg.n_cigar > eg.m_cigar) { fprintf(stderr, "thread %d: n_cigar %d > m_cigar %d\n", a->tid, eg.n_cigar, eg.m_cigar); ++a->bad; }
		if ((i & 7) == 7) { kfree(km, eg.cigar); eg.cigar = 0; eg.m_cigar = 0; }     /* the caller may free and restart from an empty buffer at any time */
	}
	kfree(km, er.cigar); kfree(km, eg.cigar);      /* both buffers came from THIS arena: kfree would corrupt the arena (or crash) otherwise */
	/* the array-of-pointers batch call with the same arena: every ez[i].cigar is krealloc'ed from km on this thread */
	if (gpu_create && gpu_extd2_batch) {
		enum { NB = 24 };
		void *ctx = gpu_create(-1);
		ksw_extz_t eb[NB], e1;
		uint8_t *tt[NB], *qq[NB]; int tls[NB], qls[NB];
		const uint8_t *tp[NB], *qp[NB];
		memset(eb, 0, sizeof eb); memset(&e1, 0, sizeof e1);
		if (!ctx) { fprintf(stderr, "thread %d: ksw2b_create failed\n", a->tid); ++a->bad; }
		for (int i = 0; i < NB; ++i) {
			tls[i] = 40 + (int)(rnd(&s) % 900); tt[i] = (uint8_t*)malloc(1024); qq[i] = (uint8_t*)malloc(1200); qls[i] = 0;
			for (int k = 0; k < tls[i]; ++k) { tt[i][k] = (uint8_t)(rnd(&s) & 3); if (rnd(&s) % 20) qq[i][qls[i]++] = (rnd(&s) % 12) ? tt[i][k] : (uint8_t)((tt[i][k] + 2) & 3); }
			if (qls[i] == 0) qq[i][qls[i]++] = 1;
			tp[i] = tt[i]; qp[i] = qq[i];
		}
		if (ctx && gpu_extd2_batch(ctx, km, NB, qls, qp, tls, tp, 5, mat, 4, 2, 24, 1, 150, 300, 0, 0, eb) != 0) { fprintf(stderr, "thread %d: batch call failed\n", a->tid); ++a->bad; }
		for (int i = 0; ctx && i < NB; ++i) {
			ksw_extd2_sse(km, qls[i], qq[i], tls[i], tt[i], 5, mat, 4, 2, 24, 1, 150, 300, 0, 0, &e1);
			++a->calls;
			if (e1.score != eb[i].score || e1.max != eb[i].max || e1.n_cigar != eb[i].n_cigar || (e1.n_cigar > 0 && memcmp(e1.cigar, eb[i].cigar, (size_t)e1.n_cigar * 4) != 0)) {
				fprintf(stderr, "thread %d batch pair %d: mismatch (score %d vs %d, n_cigar %d vs %d)\n", a->tid, i, e1.score, eb[i].score, e1.n_cigar, eb[i].n_cigar); ++a->bad;
			}
			kfree(km, eb[i].cigar);
		}
		kfree(km, e1.cigar);
		for (int i = 0; i < NB; ++i) { free(tt[i]); free(qq[i]); }
		if (ctx) gpu_destroy(ctx);
	}
	km_destroy(km);
	free(t); free(q);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: %s <libksw2_b200.so> [threads] [pairs per thread]\n", argv[0]); return 2; }
	const int nthr = argc > 2 ? atoi(argv[2]) : 6;
	if (argc > 3) n_pairs = atoi(argv[3]);
	void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	gpu_extd2 = (fn_d)dlsym(h, "ksw_extd2_sse"); gpu_extz2 = (fn_z)dlsym(h, "ksw_extz2_sse");
	if (!gpu_extd2 || !gpu_extz2) { fprintf(stderr, "symbols missing\n"); return 2; }
	gpu_create = (fn_create)dlsym(h, "ksw2b_create"); gpu_destroy = (fn_destroy)dlsym(h, "ksw2b_destroy"); gpu_extd2_batch = (fn_dbatch)dlsym(h, "ksw2b_extd2_batch");
	for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : i == j ? 2 : -4;
	pthread_t th[64]; arg_t ar[64];
	const int n = nthr < 1 ? 1 : nthr > 64 ? 64 : nthr;
	for (int i = 0; i < n; ++i) { ar[i].tid = i; ar[i].bad = 0; ar[i].calls = 0; pthread_create(&th[i], 0, worker, &ar[i]); }
	int bad = 0; long calls = 0;
	for (int i = 0; i < n; ++i) { pthread_join(th[i], 0); bad += ar[i].bad; calls += ar[i].calls; }
	printf("kalloc_interop: %d threads, %ld calls, %d mismatches\n", n, calls, bad);
	return bad ? 1 : 0;
}
