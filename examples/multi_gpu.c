/* multi_gpu.c -- one caller, every GPU of the box: a batch of mixed-length pairs with a band per pair (what a minimap2-style
 * program collects from its per-call arguments) goes through ksw2b_multi_align(); results come back in the caller's order.
 * Build:  gcc -std=c99 -Iinclude examples/multi_gpu.c -Lksw2_b200 -lksw2_b200 -Wl,-rpath,$PWD/ksw2_b200 -o multi_gpu
 * Usage:  ./multi_gpu [n_devices]   (default 1) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksw2_b200.h"

int main(int argc, char **argv)
{
	const int n_dev = argc > 1 ? atoi(argv[1]) : 1, n = 1000;
	int8_t mat[25];
	int64_t *qoff = (int64_t*)malloc(sizeof(int64_t) * (n + 1)), *toff = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
	int32_t *w = (int32_t*)malloc(sizeof(int32_t) * n);
	uint8_t *qcat, *tcat;
	ksw2b_result_t *res = (ksw2b_result_t*)malloc(sizeof(ksw2b_result_t) * n);
	const uint32_t *cigars = 0;
	ksw2b_params_t par;
	ksw2b_multi_t *set;
	unsigned s = 12345u;
	int i, j;
	for (i = 0; i < 5; ++i) for (j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? 2 : -4);
	qoff[0] = toff[0] = 0;
	for (i = 0; i < n; ++i) {                               /* lengths 100 .. 4000, band = min(500, 0.2 len + 50) */
		const int len = 100 + (int)((s = s * 1103515245u + 12345u) >> 8) % 3900;
		qoff[i + 1] = qoff[i] + len; toff[i + 1] = toff[i] + len;
		w[i] = len / 5 + 50 < 500 ? len / 5 + 50 : 500;
	}
	qcat = (uint8_t*)malloc((size_t)qoff[n]); tcat = (uint8_t*)malloc((size_t)toff[n]);
	for (j = 0; j < toff[n]; ++j) { tcat[j] = (uint8_t)(((s = s * 1103515245u + 12345u) >> 16) & 3); qcat[j] = (s >> 20) % 16 ? tcat[j] : (uint8_t)((tcat[j] + 1) & 3); }

	memset(&par, 0, sizeof par);
	par.kind = KSW2B_EXTD2; par.m = 5; par.mat = mat; par.q = 4; par.e = 2; par.q2 = 24; par.e2 = 1;
	par.w = -1; par.zdrop = 400; par.flag = KSW_EZ_RIGHT;    /* right-aligned gaps, CIGAR wanted */
	set = ksw2b_multi_create(NULL, n_dev);                   /* devices 0 .. n_dev-1; no GPU: NULL -- there is no CPU path */
	if (!set) { fprintf(stderr, "%s\n", ksw2b_last_error()); return 1; }
	if (ksw2b_multi_align(set, &par, n, qcat, qoff, tcat, toff, NULL, w, res, &cigars)) { fprintf(stderr, "%s\n", ksw2b_last_error()); return 1; }
	for (i = 0; i < 3; ++i) {
		printf("pair %d (%d bp, band %d): score %d, %d CIGAR ops:", i, (int)(toff[i + 1] - toff[i]), w[i], res[i].score, res[i].n_cigar);
		for (j = 0; j < res[i].n_cigar && j < 6; ++j) printf(" %u%c", cigars[res[i].cigar_off + j] >> 4, "MIDN"[cigars[res[i].cigar_off + j] & 0xf]);
		printf("%s\n", res[i].n_cigar > 6 ? " ..." : "");
	}
	ksw2b_multi_destroy(set);
	free(qoff); free(toff); free(w); free(qcat); free(tcat); free(res);
	return 0;
}
