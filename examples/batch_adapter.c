/* batch_adapter.c -- the caller-side change for throughput, as a maintainer of a minimap2-style program would write it:
 * the per-pair call sites stay where they are but hand their arguments to a small COLLECTOR instead of calling ksw_extd2_sse() at once;
 * when the collector is full (or the stage ends) it aligns everything with ONE ksw2b_align_ex() call -- the band travels per pair, as
 * those programs compute it per call -- and hands each call site its ksw_extz_t back.  Plain C99, no CUDA headers.
 * Build:  gcc -std=c99 -Iinclude examples/batch_adapter.c -Lksw2_b200 -lksw2_b200 -Wl,-rpath,$PWD/ksw2_b200 -o batch_adapter */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksw2_b200.h"

typedef struct {
	ksw2b_ctx_t *ctx;
	ksw2b_params_t par;              /* everything but the band is shared by the pairs of a stage */
	int64_t n, cap, qbytes, tbytes, qcap, tcap;
	uint8_t *qcat, *tcat;
	int64_t *qoff, *toff;
	int32_t *w;
	ksw_extz_t **ez;                 /* where each result goes */
} collector_t;

static void col_init(collector_t *c, const ksw2b_params_t *par)
{
	memset(c, 0, sizeof *c);
	c->ctx = ksw2b_create(-1);
	if (!c->ctx) { fprintf(stderr, "%s\n", ksw2b_last_error()); exit(1); }      /* no GPU: fail -- there is no CPU path behind this library */
	c->par = *par;
}

/* what used to be:  ksw_extd2_sse(km, qlen, q, tlen, t, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, ez);  */
static void col_add(collector_t *c, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int w, ksw_extz_t *ez)
{
	if (c->n == c->cap) {
		c->cap = c->cap ? c->cap * 2 : 1024;
		c->qoff = (int64_t*)realloc(c->qoff, sizeof(int64_t) * (size_t)(c->cap + 1)); c->toff = (int64_t*)realloc(c->toff, sizeof(int64_t) * (size_t)(c->cap + 1));
		c->w = (int32_t*)realloc(c->w, sizeof(int32_t) * (size_t)c->cap); c->ez = (ksw_extz_t**)realloc(c->ez, sizeof(ksw_extz_t*) * (size_t)c->cap);
	}
	if (c->qbytes + qlen > c->qcap) { c->qcap = (c->qbytes + qlen) * 2; c->qcat = (uint8_t*)realloc(c->qcat, (size_t)c->qcap); }
	if (c->tbytes + tlen > c->tcap) { c->tcap = (c->tbytes + tlen) * 2; c->tcat = (uint8_t*)realloc(c->tcat, (size_t)c->tcap); }
	memcpy(c->qcat + c->qbytes, q, (size_t)qlen); memcpy(c->tcat + c->tbytes, t, (size_t)tlen);
	c->qoff[c->n] = c->qbytes; c->toff[c->n] = c->tbytes;
	c->qbytes += qlen; c->tbytes += tlen;
	c->qoff[c->n + 1] = c->qbytes; c->toff[c->n + 1] = c->tbytes;
	c->w[c->n] = w; c->ez[c->n] = ez;
	++c->n;
}

/* end of the stage: one GPU batch, then every call site gets its record (ez->cigar through the caller's allocator: libc here) */
static void col_flush(collector_t *c)
{
	ksw2b_result_t *res;
	const uint32_t *cig = 0;
	int64_t i;
	if (c->n == 0) return;
	res = (ksw2b_result_t*)malloc(sizeof(ksw2b_result_t) * (size_t)c->n);
	if (ksw2b_align_ex(c->ctx, &c->par, c->n, c->qcat, c->qoff, c->tcat, c->toff, NULL, c->w, res, &cig)) { fprintf(stderr, "%s\n", ksw2b_last_error()); exit(1); }
	for (i = 0; i < c->n; ++i) {
		ksw_extz_t *ez = c->ez[i];
		ez->max = (uint32_t)res[i].max; ez->zdropped = (uint32_t)res[i].zdropped; ez->max_q = res[i].max_q; ez->max_t = res[i].max_t;
		ez->mqe = res[i].mqe; ez->mqe_t = res[i].mqe_t; ez->mte = res[i].mte; ez->mte_q = res[i].mte_q; ez->score = res[i].score; ez->reach_end = res[i].reach_end;
		ez->n_cigar = 0;
		if (res[i].n_cigar > 0 && cig) {
			if (res[i].n_cigar > ez->m_cigar) { int m = ez->m_cigar; while (m < res[i].n_cigar) m = m ? m << 1 : 4; ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)m << 2); ez->m_cigar = m; }
			memcpy(ez->cigar, cig + res[i].cigar_off, (size_t)res[i].n_cigar * 4); ez->n_cigar = res[i].n_cigar;
		}
	}
	free(res);
	c->n = 0; c->qbytes = c->tbytes = 0;
}

static void col_destroy(collector_t *c) { ksw2b_destroy(c->ctx); free(c->qcat); free(c->tcat); free(c->qoff); free(c->toff); free(c->w); free(c->ez); }

int main(void)
{
	enum { N = 2000 };
	static ksw_extz_t ez[N];
	static uint8_t t[N][600], q[N][600];
	int8_t mat[25];
	ksw2b_params_t par;
	collector_t col;
	unsigned s = 7u;
	int i, j, len[N];
	for (i = 0; i < 5; ++i) for (j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? 2 : -4);
	memset(&par, 0, sizeof par);
	par.kind = KSW2B_EXTD2; par.m = 5; par.mat = mat; par.q = 4; par.e = 2; par.q2 = 24; par.e2 = 1; par.w = -1; par.zdrop = 400; par.flag = 0;
	col_init(&col, &par);
	memset(ez, 0, sizeof ez);
	for (i = 0; i < N; ++i) {                                 /* the program's own loop over its pairs */
		len[i] = 100 + (int)((s = s * 1103515245u + 12345u) >> 8) % 500;
		for (j = 0; j < len[i]; ++j) { t[i][j] = (uint8_t)(((s = s * 1103515245u + 12345u) >> 16) & 3); q[i][j] = (s >> 20) % 10 ? t[i][j] : (uint8_t)((t[i][j] + 1) & 3); }
		col_add(&col, len[i], q[i], len[i], t[i], len[i] / 5 + 50, &ez[i]);       /* the band this program would have passed to ksw_extd2_sse */
	}
	col_flush(&col);
	for (i = 0; i < 3; ++i) printf("pair %d (%d bp): score %d, %d CIGAR ops\n", i, len[i], ez[i].score, ez[i].n_cigar);
	for (i = 0; i < N; ++i) free(ez[i].cigar);
	col_destroy(&col);
	return 0;
}
