/* ksw2_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A plain-C, one-lane-at-a-time restatement of the three Suzuki-Kasahara kernels of
 * lh3/ksw2 and of its traceback, written from the behaviour of the reference's SSE4.1
 * build.  It is the parity checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may call it.
 *
 * Parity pinning: this file is checked against (a) the unmodified reference compiled into
 * oracle/_ref/ (see oracle/Makefile) by differential fuzzing in tests/test_oracle.py and
 * (b) the golden vectors under tests/golden/ that were produced by that build.
 *
 * What is restated (reference file:line):
 *   kso_extz2  <- ksw2_extz2_sse.c:23-304   (single affine gap)
 *   kso_extd2  <- ksw2_extd2_sse.c:34-409   (two-piece affine gap)
 *   kso_exts2  <- ksw2_exts2_sse.c:33-415   (splice aware)
 *   traceback  <- ksw2.h:129-161 (ksw_backtrack, is_rot branch), ksw2.h:113-123 (push)
 *   z-drop     <- ksw2.h:191-207, reset <- ksw2.h:184-189
 *
 * The reference processes 16 int8 lanes per SSE vector and keeps its running state in one
 * zero-initialised allocation laid out as consecutive arrays.  Both facts are observable
 * (lanes outside the band but inside a 16-lane vector are computed from stale state and are
 * later consumed; score writes overshoot into the next array), so this restatement keeps
 * the same flat byte layout and the same 16-lane rounding, but evaluates every lane with
 * scalar int8 arithmetic.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ksw2.h"
#include "ksw2_oracle.h"

typedef int8_t  i8;
typedef uint8_t u8;

static inline i8 w8(int v) { return (i8)(u8)(unsigned)v; }          /* wrap to int8 */
static inline i8 smax(i8 a, i8 b) { return a > b ? a : b; }
static inline i8 smin(i8 a, i8 b) { return a < b ? a : b; }
static inline i8 umax(i8 a, i8 b) { return (u8)a > (u8)b ? a : b; }
static inline i8 umin(i8 a, i8 b) { return (u8)a < (u8)b ? a : b; }

enum { K_Z = 0, K_D = 1, K_S = 2 };

/* ---- result bookkeeping (ksw2.h:184-207) ---- */
static void ez_reset(ksw_extz_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0; ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
	ez->n_cigar = 0; ez->zdropped = 0; ez->reach_end = 0;
}

static int ez_zdrop(ksw_extz_t *ez, int32_t H, int r, int t, int zdrop, int e)
{
	if (H > (int32_t)ez->max) {
		ez->max = H; ez->max_t = t; ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) { ez->zdropped = 1; return 1; }
	}
	return 0;
}

/* ---- CIGAR (ksw2.h:113-161); memory via libc because the oracle ignores km ---- */
static void cig_push(ksw_extz_t *ez, uint32_t op, int len)
{
	if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
		if (ez->n_cigar == ez->m_cigar) {
			ez->m_cigar = ez->m_cigar ? ez->m_cigar << 1 : 4;
			ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)ez->m_cigar << 2);
		}
		ez->cigar[ez->n_cigar++] = (uint32_t)len << 4 | op;
	} else ez->cigar[ez->n_cigar - 1] += (uint32_t)len << 4;
}

static void traceback(ksw_extz_t *ez, int rev, int min_intron, const u8 *p, const int *off, const int *off_end,
                      size_t pitch, int i, int j)
{
	int state = 0, k;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		int r = i + j, force = -1;
		uint32_t d;
		if (i < off[r]) force = 2;
		if (i > off_end[r]) force = 1;
		d = force < 0 ? p[(size_t)r * pitch + (size_t)(i - off[r])] : 0;
		if (state == 0) state = d & 7;
		else if (!((d >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = d & 7;
		if (force >= 0) state = force;
		if (state == 0) { cig_push(ez, KSW_CIGAR_MATCH, 1); --i; --j; }
		else if (state == 1 || (state == 3 && min_intron <= 0)) { cig_push(ez, KSW_CIGAR_DEL, 1); --i; }
		else if (state == 3 && min_intron > 0) { cig_push(ez, KSW_CIGAR_N_SKIP, 1); --i; }
		else { cig_push(ez, KSW_CIGAR_INS, 1); --j; }
	}
	if (i >= 0) cig_push(ez, min_intron > 0 && i >= min_intron ? KSW_CIGAR_N_SKIP : KSW_CIGAR_DEL, i + 1);
	if (j >= 0) cig_push(ez, KSW_CIGAR_INS, j + 1);
	if (!rev)
		for (k = 0; k < ez->n_cigar >> 1; ++k) {
			uint32_t t = ez->cigar[k];
			ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k]; ez->cigar[ez->n_cigar - 1 - k] = t;
		}
}

/* KSW_EZ_EQX post-pass of ksw_extd2_sse (ksw2_extd2_sse.c:399-406 -> ksw_cigar2eqx, ksw2.h:163-182): M runs are split into
 * '=' / 'X' by comparing target[x+i] with query[y+i], x and y counted from the FIRST op of the array (also with
 * KSW_EZ_REV_CIGAR).  The reference's own implementation drops ksw_push_cigar's return value and corrupts the heap in
 * this commit, so this is the INTENDED behaviour -- parity unpinned for this flag. */
static void cigar_to_eqx(ksw_extz_t *ez, const u8 *query, const u8 *target)
{
	int k, i, x = 0, y = 0, n0 = ez->n_cigar;
	uint32_t *c0 = (uint32_t*)malloc((size_t)(n0 > 0 ? n0 : 1) * 4);
	memcpy(c0, ez->cigar, (size_t)n0 * 4);
	ez->n_cigar = 0;
	for (k = 0; k < n0; ++k) {
		int op = c0[k] & 0xf, len = (int)(c0[k] >> 4);
		if (op == KSW_CIGAR_MATCH) {
			for (i = 0; i < len; ++i) cig_push(ez, target[x + i] == query[y + i] ? KSW_CIGAR_EQ : KSW_CIGAR_X, 1);
			x += len; y += len;
		} else {
			cig_push(ez, (uint32_t)op, len);
			if (op == KSW_CIGAR_DEL || op == KSW_CIGAR_N_SKIP) x += len;
			else if (op == KSW_CIGAR_INS) y += len;
		}
	}
	free(c0);
}

/* ---- the engine ---- */
typedef struct {
	int kind, qlen, tlen, m, q, e, q2, e2, w, zdrop, end_bonus, flag, noncan, junc_bonus;
	const u8 *query, *target, *junc;
	const i8 *mat;
} job_t;

static __thread int64_t g_cells; /* in-band cells of the last call on this thread (SURVEY 8d cell convention) */
int64_t kso_last_cells(void) { return g_cells; }

static void engine(const job_t *J, ksw_extz_t *ez)
{
	const int kind = J->kind, qlen = J->qlen, tlen = J->tlen, m = J->m, flag = J->flag;
	int q = J->q, e = J->e, q2 = J->q2, e2 = J->e2, w = J->w;
	const int with_cigar = !(flag & KSW_EZ_SCORE_ONLY), approx = !!(flag & KSW_EZ_APPROX_MAX);
	const int right = !!(flag & KSW_EZ_RIGHT);
	int qe_h0, tlen_, qlen_, n_col_, L, r, t, last_st = -1, last_en = -1, long_thres = 0, long_diff = 0;
	int max_sc, min_sc, narr, zdrop_e;
	i8 sc_mch, sc_mis, sc_N, wild, clamp_z, init_a, init_b;
	u8 *mem, *U, *V, *X, *Y, *X2 = 0, *Y2 = 0, *DON = 0, *ACC = 0, *S, *SF, *QR, *P = 0;
	int32_t *H = 0, H0 = 0; int last_H0_t = 0;
	int *off = 0, *off_end = 0;
	size_t pitch = 0;

	g_cells = 0;
	ez_reset(ez);
	if (kind == K_Z) { if (m <= 0 || qlen <= 0 || tlen <= 0) return; }
	else { if (m <= 1 || qlen <= 0 || tlen <= 0) return; }
	if (kind == K_S && q2 <= q + e) return;
	if (kind == K_S && e == 0) return;                  /* reference divides by e (ksw2_exts2_sse.c:93): undefined there, rejected here */
	qe_h0 = q + e;                                      /* extd2: taken BEFORE the swap (ksw2_extd2_sse.c:68 vs :78) */
	if (kind == K_D && q2 + e2 < q + e) { int x; x = q; q = q2; q2 = x; x = e; e = e2; e2 = x; }
	q = (i8)q; e = (i8)e; q2 = (i8)q2; e2 = (i8)e2;

	sc_mch = J->mat[0]; sc_mis = J->mat[1];
	sc_N = J->mat[m * m - 1] == 0 ? w8(-(kind == K_D ? e2 : e)) : J->mat[m * m - 1];
	wild = (i8)(m - 1);
	clamp_z = kind == K_Z ? w8(J->mat[0] + (q + e) * 2) : J->mat[0];

	if (kind == K_S) w = tlen > qlen ? tlen : qlen;     /* exts2 has no band */
	else if (w < 0) w = tlen > qlen ? tlen : qlen;
	tlen_ = (tlen + 15) / 16; qlen_ = (qlen + 15) / 16; L = tlen_ * 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	if (kind != K_S) n_col_ = n_col_ < w + 1 ? n_col_ : w + 1;
	n_col_ = (n_col_ + 15) / 16 + 1;
	for (t = 1, max_sc = J->mat[0], min_sc = J->mat[1]; t < m * m; ++t) {
		max_sc = max_sc > J->mat[t] ? max_sc : J->mat[t];
		min_sc = min_sc < J->mat[t] ? min_sc : J->mat[t];
	}
	if (-min_sc > 2 * (q + e)) return;

	if (kind == K_D) {
		long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
		if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
		long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	} else if (kind == K_S) {
		long_thres = (q2 - q) / e - 1;
		if (q2 > q + e + long_thres * e) ++long_thres;
		long_diff = long_thres * e - (q2 - q);
	}

	/* one flat zeroed buffer; order of arrays is the reference's (extz2 :84-86, extd2 :107-110, exts2 :98-102) */
	narr = kind == K_Z ? 5 : kind == K_D ? 7 : 8;     /* state arrays up to and including s */
	mem = (u8*)calloc((size_t)(narr + 1) * L + (size_t)qlen_ * 16 + 64, 1);
	U = mem; V = U + L; X = V + L; Y = X + L;
	if (kind == K_Z) S = Y + L;
	else if (kind == K_D) { X2 = Y + L; Y2 = X2 + L; S = Y2 + L; }
	else { X2 = Y + L; DON = X2 + L; ACC = DON + L; S = ACC + L; }
	SF = S + L; QR = SF + L;
	init_a = kind == K_Z ? 0 : w8(-q - e);
	init_b = kind == K_D ? w8(-q2 - e2) : w8(-q2);
	if (kind != K_Z) { memset(U, (u8)init_a, L); memset(V, (u8)init_a, L); memset(X, (u8)init_a, L); memset(Y, (u8)init_a, L); memset(X2, (u8)init_b, L); }
	if (kind == K_D) memset(Y2, (u8)init_b, L);
	if (!approx) { H = (int32_t*)malloc((size_t)L * 4); for (t = 0; t < L; ++t) H[t] = KSW_NEG_INF; }
	if (with_cigar) {
		pitch = (size_t)n_col_ * 16;
		P = (u8*)malloc((size_t)(qlen + tlen - 1) * pitch + 16);
		off = (int*)malloc(sizeof(int) * 2 * (size_t)(qlen + tlen - 1)); off_end = off + (qlen + tlen - 1);
	}
	for (t = 0; t < qlen; ++t) QR[t] = J->query[qlen - 1 - t];
	memcpy(SF, J->target, tlen);

	if (kind == K_S && (flag & (KSW_EZ_SPLICE_FOR | KSW_EZ_SPLICE_REV))) { /* exts2 :119-171 */
		const u8 *T = J->target, *jn = J->junc;
		const int fw = !!(flag & KSW_EZ_SPLICE_FOR), rv = !!(flag & KSW_EZ_SPLICE_REV), rc = !!(flag & KSW_EZ_REV_CIGAR);
		const int semi = (flag & KSW_EZ_SPLICE_FLANK) ? -J->noncan / 2 : 0;
		/* forward CIGAR: donor GT(r)/CT(r) after t, acceptor (y)AG/(y)AC ending at t; reversed: GA(y)/CA(y), (r)TG/(r)TC */
		const int d2 = rc ? 0 : 3, df1 = rc ? 1 : 0, df2 = rc ? 3 : 2;     /* donor 2nd base; flank alternatives */
		const int a1 = rc ? 3 : 0, af1 = rc ? 0 : 1, af2 = rc ? 2 : 3;     /* acceptor 1st base; flank alternatives */
		memset(DON, (u8)w8(-J->noncan), L); memset(ACC, (u8)w8(-J->noncan), L);
		for (t = 0; t < tlen - 4; ++t) {
			int can = 0;
			if (fw && T[t+1] == 2 && T[t+2] == d2) can = 1;
			if (rv && T[t+1] == 1 && T[t+2] == d2) can = 1;
			if (can && (T[t+3] == df1 || T[t+3] == df2)) can = 2;
			if (can) DON[t] = (u8)w8(can == 2 ? 0 : semi);
		}
		if (jn) for (t = 0; t < tlen - 1; ++t) {
			int bit_f = rc ? 2 : 1, bit_r = rc ? 4 : 8;
			if ((fw && (jn[t+1] & bit_f)) || (rv && (jn[t+1] & bit_r))) DON[t] = (u8)w8((i8)DON[t] + J->junc_bonus);
		}
		for (t = 2; t < tlen; ++t) {
			int can = 0;
			if (fw && T[t-1] == a1 && T[t] == 2) can = 1;
			if (rv && T[t-1] == a1 && T[t] == 1) can = 1;
			if (can && (T[t-2] == af1 || T[t-2] == af2)) can = 2;
			if (can) ACC[t] = (u8)w8(can == 2 ? 0 : semi);
		}
		if (jn) for (t = 0; t < tlen; ++t) {
			int bit_f = rc ? 1 : 2, bit_r = rc ? 8 : 4;
			if ((fw && (jn[t] & bit_f)) || (rv && (jn[t] & bit_r))) ACC[t] = (u8)w8((i8)ACC[t] + J->junc_bonus);
		}
	}
	zdrop_e = kind == K_Z ? e : kind == K_D ? e2 : 0;

	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		i8 x1, x21 = 0, v1;
		const u8 *qrr = QR + (qlen - 1 - r);
		u8 *pr = 0;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (kind != K_S) {
			if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
			if (en > (r + w) >> 1) en = (r + w) >> 1;
			if (st > en) { ez->zdropped = 1; break; }
		}
		st0 = st; en0 = en;
		st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
		g_cells += en0 - st0 + 1;
		/* carry-in and first-row/column boundary */
		if (kind == K_Z) {
			if (st > 0) {
				if (st - 1 >= last_st && st - 1 <= last_en) { x1 = (i8)X[st - 1]; v1 = (i8)V[st - 1]; }
				else x1 = v1 = 0;
			} else { x1 = 0; v1 = r ? (i8)q : 0; }
			if (en >= r) { Y[r] = 0; U[r] = r ? (u8)q : 0; }
		} else {
			const int e_far = kind == K_D ? -e2 : 0;
			i8 bnd = w8(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : e_far);
			if (st > 0) {
				if (st - 1 >= last_st && st - 1 <= last_en) { x1 = (i8)X[st - 1]; x21 = (i8)X2[st - 1]; v1 = (i8)V[st - 1]; }
				else { x1 = init_a; x21 = init_b; v1 = init_a; }
			} else { x1 = init_a; x21 = init_b; v1 = bnd; }
			if (en >= r) { Y[r] = (u8)init_a; if (kind == K_D) Y2[r] = (u8)init_b; U[r] = (u8)bnd; }
		}
		/* score row: unaligned 16-lane chunks starting at st0, overshooting en0 (all reads precede the chunk's writes) */
		if (!(flag & KSW_EZ_GENERIC_SC)) {
			for (t = st0; t <= en0; t += 16) {
				u8 tmp[16]; int k;
				for (k = 0; k < 16; ++k) {
					u8 a = SF[t + k], b = qrr[t + k];
					tmp[k] = (a == (u8)wild || b == (u8)wild) ? (u8)sc_N : a == b ? (u8)sc_mch : (u8)sc_mis;
				}
				memcpy(S + t, tmp, 16);
			}
		} else for (t = st0; t <= en0; ++t) S[t] = (u8)J->mat[SF[t] * m + qrr[t]];
		if (with_cigar) { pr = P + (size_t)r * pitch - st; off[r] = st; off_end[r] = en; }
		/* core: every lane of the rounded range, ascending; cx/cv/cx2 hold lane t-1's OLD x/v/x2 */
		{
			i8 cx = x1, cv = v1, cx2 = x21;
			const int sx = kind == K_Z && x1 < 0, sv = kind == K_Z && v1 < 0;
			for (t = st; t <= en; ++t) {
				i8 z, a, b, a2 = 0, b2 = 0, a2a = 0, xt1 = cx, vt1 = cv, x2t1 = cx2, ut, zq;
				u8 d = 0;
				if (t > st && t < st + 4) { if (sx) xt1 = -1; if (sv) vt1 = -1; } /* extz2 sign-extension quirk */
				cx = (i8)X[t]; cv = (i8)V[t]; if (kind != K_Z) cx2 = (i8)X2[t];
				ut = (i8)U[t];
				a = w8(xt1 + vt1); b = w8((i8)Y[t] + ut);
				if (kind == K_Z) {
					z = w8((i8)S[t] + (q + e) * 2);
					if (!with_cigar) z = smax(z, a);
					else if (!right) { d = a > z ? 1 : 0; z = smax(z, a); if (b > z) d = 2; }
					else { d = z > a ? 0 : 1; z = smax(z, a); d = z > b ? d : 2; }
					z = umax(z, b); z = umin(z, clamp_z);
					U[t] = (u8)w8(z - vt1); V[t] = (u8)w8(z - ut);
					zq = w8(z - q); a = w8(a - zq); b = w8(b - zq);
					if (!with_cigar || !right) {
						X[t] = a > 0 ? (u8)a : 0; Y[t] = b > 0 ? (u8)b : 0;
						if (a > 0) d |= 0x08;
						if (b > 0) d |= 0x10;
					} else {
						X[t] = 0 > a ? 0 : (u8)a; Y[t] = 0 > b ? 0 : (u8)b;
						if (!(0 > a)) d |= 0x08;
						if (!(0 > b)) d |= 0x10;
					}
				} else {
					i8 zq2, don = 0;
					z = (i8)S[t];
					a2 = w8(x2t1 + vt1);
					if (kind == K_D) b2 = w8((i8)Y2[t] + ut);
					else { a2a = w8(a2 + (i8)ACC[t]); don = (i8)DON[t]; }
					if (!right) {
						if (a > z) d = 1;
						z = smax(z, a);
						if (b > z) d = 2;
						z = smax(z, b);
						if (kind == K_D) {
							if (a2 > z) d = 3;
							z = smax(z, a2);
							if (b2 > z) d = 4;
							z = smax(z, b2);
						} else { if (a2a > z) d = 3; z = smax(z, a2a); }
					} else {
						d = z > a ? 0 : 1; z = smax(z, a);
						d = z > b ? d : 2; z = smax(z, b);
						if (kind == K_D) {
							d = z > a2 ? d : 3; z = smax(z, a2);
							d = z > b2 ? d : 4; z = smax(z, b2);
						} else { d = z > a2a ? d : 3; z = smax(z, a2a); }
					}
					if (kind == K_D) z = smin(z, clamp_z);
					U[t] = (u8)w8(z - vt1); V[t] = (u8)w8(z - ut);
					zq = w8(z - q); zq2 = w8(z - q2);
					a = w8(a - zq); b = w8(b - zq); a2 = w8(a2 - zq2); if (kind == K_D) b2 = w8(b2 - zq2);
					/* x = max(a,0) - (q+e) etc.; left/right only differ in the tie a==0 for the flag bits */
					X[t] = (u8)w8((a > 0 ? a : 0) - (q + e)); Y[t] = (u8)w8((b > 0 ? b : 0) - (q + e));
					if (!right) { if (a > 0) d |= 0x08; if (b > 0) d |= 0x10; }
					else { if (!(0 > a)) d |= 0x08; if (!(0 > b)) d |= 0x10; }
					if (kind == K_D) {
						X2[t] = (u8)w8((a2 > 0 ? a2 : 0) - (q2 + e2)); Y2[t] = (u8)w8((b2 > 0 ? b2 : 0) - (q2 + e2));
						if (!right) { if (a2 > 0) d |= 0x20; if (b2 > 0) d |= 0x40; }
						else { if (!(0 > a2)) d |= 0x20; if (!(0 > b2)) d |= 0x40; }
					} else {
						X2[t] = (u8)w8(smax(a2, don) - q2);
						if (!right) { if (a2 > don) d |= 0x20; }
						else { if (!(don > a2)) d |= 0x20; }
					}
				}
				if (with_cigar) pr[t] = d;
			}
		}
		if (!approx) {
			int32_t max_H, max_t;
			const int qe_sub = kind == K_Z ? q + e : 0;     /* extz2 keeps unsigned u,v and subtracts q+e */
#define UV(arr, i) (kind == K_Z ? (int32_t)(arr)[i] : (int32_t)(i8)(arr)[i])
			if (r > 0) {
				int32_t HH[4], tt[4]; int en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + UV(U, en0) - qe_sub : H[en0] + UV(V, en0) - qe_sub;
				max_t = en0;
				for (i = 0; i < 4; ++i) { HH[i] = max_H; tt[i] = max_t; }
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += UV(V, t + i) - qe_sub;
						if (H[t + i] > HH[i]) { HH[i] = H[t + i]; tt[i] = t; }
					}
				for (i = 0; i < 4; ++i) if (max_H < HH[i]) { max_H = HH[i]; max_t = tt[i] + i; }
				for (; t < en0; ++t) { H[t] += UV(V, t) - qe_sub; if (H[t] > max_H) { max_H = H[t]; max_t = t; } }
			} else {
				H[0] = UV(V, 0) - (kind == K_Z ? 2 * (q + e) : kind == K_D ? qe_h0 : q + e);
				max_H = H[0]; max_t = 0;
			}
			if (en0 == tlen - 1 && H[en0] > ez->mte) { ez->mte = H[en0]; ez->mte_q = r - en; }
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) { ez->mqe = H[st0]; ez->mqe_t = st0; }
			if (ez_zdrop(ez, max_H, r, max_t, J->zdrop, zdrop_e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else {
			const int qe_sub = kind == K_Z ? q + e : 0;
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = UV(V, last_H0_t) - qe_sub, d1 = UV(U, last_H0_t + 1) - qe_sub;
					if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += UV(V, last_H0_t) - qe_sub;
				else { ++last_H0_t; H0 += UV(U, last_H0_t) - qe_sub; }
				if (kind == K_Z && (flag & KSW_EZ_APPROX_DROP) && ez_zdrop(ez, H0, r, last_H0_t, J->zdrop, zdrop_e)) break;
			} else { H0 = UV(V, 0) - (kind == K_Z ? 2 * (q + e) : kind == K_D ? qe_h0 : q + e); last_H0_t = 0; }
			if (kind != K_Z && (flag & KSW_EZ_APPROX_DROP) && ez_zdrop(ez, H0, r, last_H0_t, J->zdrop, zdrop_e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
#undef UV
		last_st = st; last_en = en;
	}
	free(mem); free(H);
	if (with_cigar) {
		const int rev = !!(flag & KSW_EZ_REV_CIGAR), mil = kind == K_S ? long_thres : 0;
		if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY))
			traceback(ez, rev, mil, P, off, off_end, pitch, tlen - 1, qlen - 1);
		else if (kind != K_S && !ez->zdropped && (flag & KSW_EZ_EXTZ_ONLY) && ez->mqe + J->end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			traceback(ez, rev, mil, P, off, off_end, pitch, ez->mqe_t, qlen - 1);
		} else if (ez->max_t >= 0 && ez->max_q >= 0)
			traceback(ez, rev, mil, P, off, off_end, pitch, ez->max_t, ez->max_q);
		if (kind == K_D && (flag & KSW_EZ_EQX)) cigar_to_eqx(ez, J->query, J->target);
		free(P); free(off);
	}
}

void kso_extz2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	job_t J; (void)km; memset(&J, 0, sizeof J);
	J.kind = K_Z; J.qlen = qlen; J.tlen = tlen; J.query = query; J.target = target; J.m = m; J.mat = mat;
	J.q = q; J.e = e; J.w = w; J.zdrop = zdrop; J.end_bonus = end_bonus; J.flag = flag;
	engine(&J, ez);
}

void kso_extd2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	job_t J; (void)km; memset(&J, 0, sizeof J);
	J.kind = K_D; J.qlen = qlen; J.tlen = tlen; J.query = query; J.target = target; J.m = m; J.mat = mat;
	J.q = q; J.e = e; J.q2 = q2; J.e2 = e2; J.w = w; J.zdrop = zdrop; J.end_bonus = end_bonus; J.flag = flag;
	engine(&J, ez);
}

void kso_exts2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int8_t q2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez)
{
	job_t J; (void)km; memset(&J, 0, sizeof J);
	J.kind = K_S; J.qlen = qlen; J.tlen = tlen; J.query = query; J.target = target; J.m = m; J.mat = mat;
	J.q = q; J.e = e; J.q2 = q2; J.noncan = noncan; J.zdrop = zdrop; J.junc_bonus = junc_bonus; J.flag = flag; J.junc = junc;
	engine(&J, ez);
}

/* ---- the two ROW-WISE scalar entry points of the reference API --------------------------------------------
 *   kso_extz <- ksw2_extz.c:6-135   (single affine gap, int32, one target row at a time)
 *   kso_extd <- ksw2_extd.c:6-175   (two-piece affine gap)
 * Not the Suzuki-Kasahara formulation: cells outside the band are -inf (ksw2_extz.c:35,43-44), the maximum and the
 * Z-drop test are per ROW (:116-122, ksw_apply_zdrop with is_rot == 0), wildcard scores come from `mat`, there is no
 * end_bonus / reach_end.  Traceback = ksw_backtrack with is_rot == 0 and off_end == NULL (ksw2.h:129-161).
 * Outside the parity domain (undefined in the reference, deterministic here): a row whose band starts beyond the query end
 * (the reference writes past eh[], :113) and a traceback cell right of the band (the reference reads unwritten heap). */
typedef struct { int32_t *h, *e, *e2; } rowstate_t;

static void rows_traceback(ksw_extz_t *ez, int rev, const u8 *z, size_t n_col, int w, int qlen, int i, int j)
{
	int state = 0, k;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		const int st = i > w ? i - w : 0, en = i + w < qlen - 1 ? i + w : qlen - 1;
		int force = -1;
		uint32_t d;
		if (j < st) force = 2;                 /* ksw2.h:141 */
		else if (j > en) force = 1;            /* not in the reference (off_end == NULL): see the note above */
		d = force < 0 ? z[(size_t)i * n_col + (size_t)(j - st)] : 0;
		if (state == 0) state = d & 7;
		else if (!((d >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = d & 7;
		if (force >= 0) state = force;
		if (state == 0) { cig_push(ez, KSW_CIGAR_MATCH, 1); --i; --j; }
		else if (state == 1 || state == 3) { cig_push(ez, KSW_CIGAR_DEL, 1); --i; }
		else { cig_push(ez, KSW_CIGAR_INS, 1); --j; }
	}
	if (i >= 0) cig_push(ez, KSW_CIGAR_DEL, i + 1);
	if (j >= 0) cig_push(ez, KSW_CIGAR_INS, j + 1);
	if (!rev)
		for (k = 0; k < ez->n_cigar >> 1; ++k) {
			uint32_t t = ez->cigar[k];
			ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k]; ez->cigar[ez->n_cigar - 1 - k] = t;
		}
}

/* a "takes over" b?  left-aligned: ties keep the earlier candidate; right-aligned: ties go to the later one */
static inline int beats(int32_t cand, int32_t cur, int right) { return right ? cand >= cur : cand > cur; }

static void rows_engine(int dual /* 0 ksw_extz, 1 ksw_extd, 2 ksw_gg */, int qlen, const u8 *query, int tlen, const u8 *target, int m, const i8 *mat,
                        int go, int ge, int go2, int ge2, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	const int with_cigar = !(flag & KSW_EZ_SCORE_ONLY), right = with_cigar && (flag & KSW_EZ_RIGHT);
	const int oe = go + ge, oe2 = go2 + ge2;
	const int gg = dual == 2;
	rowstate_t S;
	u8 *z = 0;
	size_t n_col;
	int i, j, max_j = 0;

	if (gg) dual = 0;                                                          /* ksw_gg: ksw_extz's arithmetic, left-aligned, no extension logic */
	g_cells = 0;
	ez_reset(ez);
	if (w < 0) w = tlen > qlen ? tlen : qlen;                                  /* :15 */
	n_col = (size_t)(qlen < 2 * w + 1 ? qlen : 2 * w + 1);                     /* :16 */
	S.h = (int32_t*)calloc((size_t)qlen + 1, 4); S.e = (int32_t*)calloc((size_t)qlen + 1, 4); S.e2 = (int32_t*)calloc((size_t)qlen + 1, 4);
	if (with_cigar) z = (u8*)calloc(n_col * (size_t)tlen + 1, 1);
	/* row "-1": costs of a leading insertion of length j (:32-35 / ksw2_extd.c:33-41) */
	for (j = 0; j <= qlen; ++j) {
		if (j > w && j > 0) { S.h[j] = S.e[j] = KSW_NEG_INF; S.e2[j] = dual ? KSW_NEG_INF : 0; continue; }
		if (j == 0) { S.h[0] = 0; S.e[0] = -2 * oe; S.e2[0] = dual ? -2 * oe2 : 0; }
		else if (!dual) { S.h[j] = -(oe + ge * (j - 1)); S.e[j] = -(2 * oe + ge * j); }
		else {
			const int c1 = go + ge * j, c2 = go2 + ge2 * j, d1 = oe + ge * j, d2 = oe2 + ge2 * j;
			const int best = d1 < d2 ? -d1 : -d2;
			S.h[j] = c1 < c2 ? -c1 : -c2; S.e[j] = best - oe; S.e2[j] = best - oe2;
		}
	}
	for (i = 0; i < tlen; ++i) {
		const i8 *srow = mat + (size_t)target[i] * m;
		const int st = i > w ? i - w : 0, en = i + w < qlen - 1 ? i + w : qlen - 1;
		int32_t rowmax = KSW_NEG_INF, left_h, f, f2 = 0;
		/* column "-1" of this row: a leading deletion of length i+1, or -inf once the band has left column 0 (:44-45) */
		if (st > 0) left_h = f = f2 = KSW_NEG_INF;
		else if (!dual) { left_h = -(oe + ge * i); f = -(2 * oe + ge * i); }
		else {
			const int d1 = oe + ge * i, d2 = oe2 + ge2 * i, best = d1 < d2 ? -d1 : -d2;
			left_h = best; f = best - oe; f2 = best - oe2;
		}
		for (j = st; j <= en; ++j) {
			int32_t diag = S.h[j], e = S.e[j], e2 = S.e2[j], h, hop;
			int d = 0;
			S.h[j] = left_h;
			h = diag + srow[query[j]];
			/* candidates in the reference's order H, E, F, E2, F2 (:71-74 left, :98-101 right) */
			if (right ? !(h > e) : !(h >= e)) { h = e; d = 1; }
			if (right ? !(h > f) : !(h >= f)) { h = f; d = 2; }
			if (dual) {
				if (right ? !(h > e2) : !(h >= e2)) { h = e2; d = 3; }
				if (right ? !(h > f2) : !(h >= f2)) { h = f2; d = 4; }
			}
			left_h = h;
			/* row maximum: the LAST maximal column, except right-aligned ksw_extz which keeps the first (ksw2_extz.c:103-104) */
			if (right && !dual ? h > rowmax : h >= rowmax) { max_j = j; rowmax = h; }
			hop = h - oe;
			e -= ge; if (beats(e, hop, right)) d |= 0x08; else e = hop;
			f -= ge; if (beats(f, hop, right)) d |= 0x10; else f = hop;
			S.e[j] = e;
			if (dual) {
				const int32_t hop2 = h - oe2;
				e2 -= ge2; if (beats(e2, hop2, right)) d |= 0x20; else e2 = hop2;
				f2 -= ge2; if (beats(f2, hop2, right)) d |= 0x40; else f2 = hop2;
				S.e2[j] = e2;
			}
			if (with_cigar) z[(size_t)i * n_col + (size_t)(j - st)] = (u8)d;
			++g_cells;
		}
		if (gg) j = en + 1;                                                    /* ksw_gg writes eh[en] with its EXCLUSIVE en (ksw2_gg.c:92): always inside eh[] */
		if (j <= qlen) { S.h[j] = left_h; S.e[j] = KSW_NEG_INF; }              /* :113 (e2 of that column is left alone) */
		if (en == qlen - 1 && S.h[qlen] > ez->mqe) { ez->mqe = S.h[qlen]; ez->mqe_t = i; }
		if (i == tlen - 1) { ez->mte = rowmax; ez->mte_q = max_j; }
		if (ez_zdrop(ez, rowmax, i + max_j, i, zdrop, dual ? ge2 : ge)) break;  /* is_rot == 0: r = i + max_j, t = i */
		if (i == tlen - 1 && en == qlen - 1) ez->score = S.h[qlen];
	}
	if (gg) { const int32_t sc = S.h[qlen]; ez_reset(ez); ez->score = sc; }    /* ksw2_gg.c:96: whatever eh[qlen].h holds; nothing else is reported */
	free(S.h); free(S.e); free(S.e2);
	if (with_cigar) {
		const int rev = !!(flag & KSW_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY)) rows_traceback(ez, rev, z, n_col, w, qlen, tlen - 1, qlen - 1);
		else if (ez->max_t >= 0 && ez->max_q >= 0) rows_traceback(ez, rev, z, n_col, w, qlen, ez->max_t, ez->max_q);
		free(z);
	}
}

void kso_extz(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t gapo, int8_t gape, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	(void)km;
	rows_engine(0, qlen, query, tlen, target, m, mat, gapo, gape, 0, 0, w, zdrop, flag, ez);
}

void kso_extd(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	(void)km;
	rows_engine(1, qlen, query, tlen, target, m, mat, gapo, gape, gapo2, gape2, w, zdrop, flag, ez);
}

/* ---- kso_extf2 <- ksw2_extf2_sse.c:11-98: linear-gap u/v recurrence with X-drop, score only ----------
 * Lane-at-a-time restatement of the SSE4.1 build.  Observable details kept: ONE zeroed allocation laid out as u | v | s | sf
 * (target copy) | qr (reversed query) (:26-28), so the unaligned 16-lane score chunks that start at st0 (:52-62) read past sf
 * into qr and write past s into the first bytes of sf; lanes are evaluated in whole 16-lane vectors st..en (:63-84); the
 * tracked cell H0 reads u/v as UNSIGNED bytes (:85-96). */
void kso_extf2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t mch, int8_t mis, int8_t e, int w, int xdrop, ksw_extz_t *ez)
{
	int r, t, last_st = -1, last_en = -1, last_t = 0;
	int32_t H0 = 0;
	const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16, L = tlen_ * 16;
	const i8 sc_mis = mis < 0 ? mis : (i8)-mis, e2 = w8(e * 2);
	u8 *mem, *U, *V, *S, *SF, *QR;
	(void)km;
	g_cells = 0;
	ez_reset(ez);
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	mem = (u8*)calloc((size_t)(tlen_ * 4 + qlen_ + 2) * 16, 1);
	U = mem; V = U + L; S = V + L; SF = S + L; QR = SF + L;
	for (t = 0; t < qlen; ++t) QR[t] = query[qlen - 1 - t];
	memcpy(SF, target, (size_t)tlen);
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		const u8 *qrr = QR + (qlen - 1 - r);
		u8 carry;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
		if (en > ((r + w) >> 1)) en = (r + w) >> 1;
		if (st > en) break;
		st0 = st; en0 = en;
		st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
		carry = (st > 0 && st - 1 >= last_st && st - 1 <= last_en) ? V[st - 1] : 0;
		if (en >= r) U[r] = 0;
		for (t = st0; t <= en0; t += 16) {                 /* score chunks: 16 lanes each, unaligned, overshooting en0 */
			int k;
			u8 tmp[16];
			for (k = 0; k < 16; ++k) tmp[k] = (u8)(SF[t + k] == qrr[t + k] ? mch : sc_mis);   /* loads first (:53-55), then the store (:61) */
			memcpy(S + t, tmp, 16);
		}
		for (t = st; t <= en; ++t) {
			const i8 vt1 = (i8)carry, ut = (i8)U[t];
			i8 z = w8((i8)S[t] + e2);
			carry = V[t];
			z = smax(z, vt1);                               /* SSE4.1: signed max (:72) */
			z = umax(z, ut);                                /* unsigned max (:77) */
			U[t] = (u8)w8(z - vt1); V[t] = (u8)w8(z - ut);
		}
		g_cells += en0 - st0 + 1;
		if (r > 0) {
			if (last_t >= st0 && last_t <= en0 && last_t + 1 >= st0 && last_t + 1 <= en0) {
				const int32_t d0 = V[last_t] - e, d1 = U[last_t + 1] - e;
				if (d0 > d1) H0 += d0; else { H0 += d1; ++last_t; }
			} else if (last_t >= st0 && last_t <= en0) H0 += V[last_t] - e;
			else { ++last_t; H0 += U[last_t] - e; }
			if (H0 > (int32_t)ez->max) { ez->max = H0; ez->max_t = last_t; ez->max_q = r - last_t; }
			else if (xdrop >= 0 && (int32_t)ez->max - H0 > xdrop) break;
		} else { H0 = V[0] - e - e; last_t = 0; }
		last_st = st; last_en = en;
	}
	if (r == qlen + tlen - 1) ez->score = H0; else ez->zdropped = 1;
	free(mem);
}

/* kso_gg <- ksw2_gg.c:6-102: global alignment, row-wise; score + (optionally) CIGAR through the caller's three pointers */
int kso_gg(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int w,
           int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	ksw_extz_t ez;
	const int with = m_cigar_ && n_cigar_ && cigar_;
	(void)km;
	memset(&ez, 0, sizeof ez);
	if (with) { ez.cigar = *cigar_; ez.m_cigar = *m_cigar_; }
	rows_engine(2, qlen, query, tlen, target, m, mat, gapo, gape, 0, 0, w, -1, with ? 0 : KSW_EZ_SCORE_ONLY, &ez);
	if (with) { *cigar_ = ez.cigar; *m_cigar_ = ez.m_cigar; *n_cigar_ = ez.n_cigar; }
	return ez.score;
}

/* ---- the two anti-diagonal GLOBAL alignment entry points ------------------------------------------------
 *   kso_gg2     <- ksw2_gg2.c:4-114      scalar int8, exact band [st, en], signed compares, boundary tests by band geometry
 *   kso_gg2_sse <- ksw2_gg2_sse.c:11-126 16-lane vectors st..en (rounded), score row only on [st0, en0], unsigned max for b,
 *                                         NO clamp of z, stale-neighbour test by last_st/last_en, H0 from unsigned bytes
 * Both keep the direction bytes in one flat allocation with the reference's row pitch and hand it to ksw_backtrack with
 * is_rot == 1 and off_end == NULL (ksw2.h:129-161): a cell right of the band is read from wherever the flat index lands.
 * gg2 zeroes that allocation (kcalloc, :20); gg2_sse does not (kmalloc, :37) -- zeroed here, outside the parity domain. */
static void rot_traceback_flat(ksw_extz_t *ez, const u8 *p, size_t psize, const int *off, size_t n_col, int i, int j)
{
	int state = 0, k;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		const int r = i + j;
		int force = -1;
		uint32_t d;
		size_t idx;
		if (i < off[r]) force = 2;
		idx = (size_t)r * n_col + (size_t)(i - off[r]);
		d = force < 0 ? (idx < psize ? p[idx] : 0) : 0;
		if (state == 0) state = d & 7;
		else if (!((d >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = d & 7;
		if (force >= 0) state = force;
		if (state == 0) { cig_push(ez, KSW_CIGAR_MATCH, 1); --i; --j; }
		else if (state == 1 || state == 3) { cig_push(ez, KSW_CIGAR_DEL, 1); --i; }
		else { cig_push(ez, KSW_CIGAR_INS, 1); --j; }
	}
	if (i >= 0) cig_push(ez, KSW_CIGAR_DEL, i + 1);
	if (j >= 0) cig_push(ez, KSW_CIGAR_INS, j + 1);
	for (k = 0; k < ez->n_cigar >> 1; ++k) {
		uint32_t t = ez->cigar[k];
		ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k]; ez->cigar[ez->n_cigar - 1 - k] = t;
	}
}

static int gg2_engine(int sse, int qlen, const u8 *query, int tlen, const u8 *target, int m, const i8 *mat, int q, int e, int w, int with, ksw_extz_t *ez)
{
	const int qe = q + e, tlen_ = (tlen + 15) / 16, L = tlen_ * 16 + 16;
	const i8 qe2 = w8(qe * 2);
	int r, t, n_col, *off = 0, H0 = 0, last_t = 0, last_st = -1, last_en = -1;
	i8 *U, *V, *X, *Y, *S;
	u8 *P = 0;
	size_t pitch, psize = 0;
	g_cells = 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	n_col = w + 1 < tlen ? w + 1 : tlen;
	pitch = sse ? (size_t)((n_col + 15) / 16 + 1) * 16 : (size_t)n_col;
	U = (i8*)calloc((size_t)L * 5, 1); V = U + L; X = V + L; Y = X + L; S = Y + L;
	if (with) {
		psize = (size_t)(qlen + tlen) * pitch + 16;
		P = (u8*)calloc(psize, 1);
		off = (int*)calloc((size_t)(qlen + tlen), sizeof(int));
	}
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		i8 x1, v1;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
		if (en > ((r + w) >> 1)) en = (r + w) >> 1;
		st0 = st; en0 = en;
		if (sse) { st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1; }
		if (with) off[r] = st;
		if (!sse) {                                                            /* ksw2_gg2.c:36-43 */
			if (st != 0) { if (r > st + st + w - 1) x1 = v1 = 0; else { x1 = X[st - 1]; v1 = V[st - 1]; } }
			else { x1 = 0; v1 = (i8)(r ? q : 0); }
			if (en != r) { if (r < en + en - w - 1) Y[en] = U[en] = 0; }
			else { Y[r] = 0; U[r] = (i8)(r ? q : 0); }
		} else {                                                               /* ksw2_gg2_sse.c:54-60 */
			if (st > 0) { if (st - 1 >= last_st && st - 1 <= last_en) { x1 = X[st - 1]; v1 = V[st - 1]; } else x1 = v1 = 0; }
			else { x1 = 0; v1 = (i8)(r ? q : 0); }
			if (en >= r) { Y[r] = 0; U[r] = (i8)(r ? q : 0); }
		}
		for (t = st0; t <= en0; ++t) S[t] = mat[target[t] * m + query[qlen - 1 - (t + qlen - 1 - r)]];
		for (t = st; t <= en; ++t) {
			i8 z = w8(S[t] + qe2), a = w8(x1 + v1), b = w8(Y[t] + U[t]), u1;
			int d = a > z ? 1 : 0;
			z = smax(z, a);
			d = b > z ? 2 : d;
			z = sse ? umax(z, b) : smax(z, b);
			u1 = U[t]; U[t] = w8(z - v1); v1 = V[t]; V[t] = w8(z - u1);
			z = w8(z - q); a = w8(a - z); b = w8(b - z);
			x1 = X[t];
			if (a > 0) d |= 0x08;
			X[t] = a > 0 ? a : 0;
			if (b > 0) d |= 0x10;
			Y[t] = b > 0 ? b : 0;
			if (with) P[(size_t)r * pitch + (size_t)(t - st)] = (u8)d;
		}
		g_cells += en0 - st0 + 1;
		if (r > 0) {
			if (last_t >= st0 && last_t <= en0) H0 += (sse ? (int)(u8)V[last_t] : (int)V[last_t]) - qe;
			else { ++last_t; H0 += (sse ? (int)(u8)U[last_t] : (int)U[last_t]) - qe; }
		} else { H0 = (sse ? (int)(u8)V[0] : (int)V[0]) - 2 * qe; last_t = 0; }
		last_st = st; last_en = en;
	}
	free(U);
	ez_reset(ez);
	ez->score = H0;
	if (with) { rot_traceback_flat(ez, P, psize, off, pitch, tlen - 1, qlen - 1); free(P); free(off); }
	return H0;
}

static int gg2_call(int sse, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                    int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	ksw_extz_t ez;
	const int with = m_cigar_ && n_cigar_ && cigar_;
	int sc;
	memset(&ez, 0, sizeof ez);
	if (with) { ez.cigar = *cigar_; ez.m_cigar = *m_cigar_; }
	sc = gg2_engine(sse, qlen, query, tlen, target, m, mat, q, e, w, with, &ez);
	if (with) { *cigar_ = ez.cigar; *m_cigar_ = ez.m_cigar; *n_cigar_ = ez.n_cigar; }
	return sc;
}
int kso_gg2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
            int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{ (void)km; return gg2_call(0, qlen, query, tlen, target, m, mat, q, e, w, m_cigar_, n_cigar_, cigar_); }
int kso_gg2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{ (void)km; return gg2_call(1, qlen, query, tlen, target, m, mat, q, e, w, m_cigar_, n_cigar_, cigar_); }
