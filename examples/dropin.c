/* dropin.c -- what a ksw2 / minimap2-style caller looks like against this library (plain C99, no CUDA headers).
 * Build:  gcc -std=c99 -Iinclude examples/dropin.c -Lksw2_b200 -lksw2_b200 -Wl,-rpath,$PWD/ksw2_b200 -o dropin
 * (1) the reference's own call, unchanged (README.md:60-75 of lh3/ksw2); (2) the same work as one batch. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksw2.h"
#include "ksw2_b200.h"

static void encode(const char *s, uint8_t *out, int n)
{
	int i;
	for (i = 0; i < n; ++i) out[i] = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : s[i] == 'T' ? 3 : 4;
}

int main(void)
{
	const char *ts = "ATAGCTAGCTAGCAT", *qs = "AGCTACCGCAT";
	const int tl = (int)strlen(ts), ql = (int)strlen(qs);
	int8_t a = 1, b = -2, mat[25];
	uint8_t t[64], q[64];
	ksw_extz_t ez;
	int i, j;
	for (i = 0; i < 5; ++i) for (j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? a : b);
	encode(ts, t, tl); encode(qs, q, ql);

	/* (1) one pair per call, exactly as with the reference */
	memset(&ez, 0, sizeof ez);
	ksw_extz2_sse(0, ql, q, tl, t, 5, mat, 2, 1, -1, -1, 0, 0, &ez);
	printf("score %d, %d CIGAR ops:", ez.score, ez.n_cigar);
	for (i = 0; i < ez.n_cigar; ++i) printf(" %u%c", ez.cigar[i] >> 4, "MIDN"[ez.cigar[i] & 0xf]);
	printf("\n");
	free(ez.cigar);                                  /* km == NULL: libc realloc / free, like the reference without kalloc */

	/* (2) many pairs per call */
	{
		ksw2b_ctx_t *ctx = ksw2b_create(-1);
		const int n = 4;
		int qlen[4], tlen[4];
		const uint8_t *qp[4], *tp[4];
		ksw_extz_t ezs[4];
		if (!ctx) { fprintf(stderr, "%s\n", ksw2b_last_error()); return 1; }
		for (i = 0; i < n; ++i) { qlen[i] = ql; tlen[i] = tl; qp[i] = q; tp[i] = t; }
		memset(ezs, 0, sizeof ezs);
		if (ksw2b_extz2_batch(ctx, 0, n, qlen, qp, tlen, tp, 5, mat, 2, 1, -1, -1, 0, 0, ezs)) { fprintf(stderr, "%s\n", ksw2b_last_error()); return 1; }
		for (i = 0; i < n; ++i) { printf("pair %d: score %d\n", i, ezs[i].score); free(ezs[i].cigar); }
		ksw2b_destroy(ctx);
	}
	return 0;
}
