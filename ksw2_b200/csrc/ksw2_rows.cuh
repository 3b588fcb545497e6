// ksw2_rows.cuh -- the two scalar entry points of the reference API, ksw_extz() and ksw_extd()
// (ksw2_extz.c:6-135, ksw2_extd.c:6-175), as a GPU code path: one thread evaluates one pair row by row.
//
// These are NOT the Suzuki-Kasahara kernels: they are the row-wise (Green / AE86) int32 formulation with its
// own observable semantics -- -inf outside the band (ksw2_extz.c:35,43-44), one max / Z-drop test per ROW
// (:116-122), wildcard scores taken from `mat`, no end_bonus -- so they get their own small kernel instead
// of a mode of the tile engine.  They exist so that a caller linking the whole ksw2.h API finds every symbol
// served by the GPU (north_star names ksw_extz / ksw_extd); throughput comes from batching many pairs, one
// per thread, not from intra-pair parallelism.
//
// Per-pair scratch (global memory): eh[qlen+1] = {h, e, e2} int32 (ksw2_extz.c:4,18 / ksw2_extd.c:4,19), element stride `es`
// (device: 32, the rows of a warp's 32 pairs interleaved word by word so that every access of the warp is one 128-byte line;
// host simulator: 1); direction bytes z[tlen][n_col] with n_col = min(qlen, 2w+1) (:16,20); off[i] = st(i) is recomputed.
// Traceback = ksw_backtrack with is_rot == 0 (ksw2.h:129-161).
#pragma once
#include "ksw2_pair.cuh"

enum { KS_ROWZ = 3, KS_ROWD = 4, KS_ROWG = 6 };   // row-wise entry points: ksw_extz, ksw_extd, ksw_gg (= ksw_extz's arithmetic, global, ksw2_gg.c:6-102)

struct KsRowsParams {                    // what ksw_extz / ksw_extd take besides the sequences
	int kind, m, gapo, gape, gapo2, gape2, w, zdrop, flag;
	const int8_t *mat;                   // m*m scores, device pointer
};

KS_HD int ks_rows_w(const KsRowsParams &P, int qlen, int tlen) { return P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w; }   // :15
KS_HD long long ks_rows_ncol(const KsRowsParams &P, int qlen, int tlen) { const long long w = ks_rows_w(P, qlen, tlen); return qlen < 2 * w + 1 ? qlen : 2 * w + 1; }   // :16
KS_HD size_t ks_rows_eh_words(int qlen) { return (size_t)(qlen + 1) * 3; }   // int32 words per pair
KS_HD size_t ks_rows_z_bytes(const KsRowsParams &P, int qlen, int tlen) { return (size_t)ks_rows_ncol(P, qlen, tlen) * (size_t)tlen; }

// ksw_apply_zdrop with is_rot == 0 (ksw2.h:191-207): a = i, b = max_j
KS_HD bool ks_rows_zdrop(KsEz &ez, int H, int i, int j, int zdrop, int e)
{
	const int r = i + j, t = i;
	if (H > ez.max) { ez.max = H; ez.max_t = t; ez.max_q = r - t; }
	else if (t >= ez.max_t && r - t >= ez.max_q) {
		const int tl = t - ez.max_t, ql = (r - t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez.max - H > zdrop + l * e) { ez.zdropped = 1; return true; }
	}
	return false;
}

// DP fill.  eh: 3 int32 per query column (h, e, e2), z: direction bytes (CIGAR runs only).
KS_HD void ks_rows_fill(const KsRowsParams &P, const uint8_t *query, int qlen, const uint8_t *target, int tlen, int32_t *eh, int es, uint8_t *z, KsEz &ez)
{
	const bool dual = P.kind == KS_ROWD, gg = P.kind == KS_ROWG, cig = !(P.flag & KSF_SCORE_ONLY);
	const bool right = cig && (P.flag & KSF_RIGHT) != 0;      // the score-only loop is the left-aligned arithmetic whatever the flag says (:47)
	const int gapo = P.gapo, gape = P.gape, gapo2 = P.gapo2, gape2 = P.gape2, gapoe = gapo + gape, gapoe2 = gapo2 + gape2;
	const int w = ks_rows_w(P, qlen, tlen);
	const long long n_col = ks_rows_ncol(P, qlen, tlen);
	int max_j = 0;
	ks_ez_reset(ez); ez.n_diag = tlen;
#define EH(j) eh[3 * (size_t)(j) * es]
#define EE(j) eh[(3 * (size_t)(j) + 1) * es]
#define EE2(j) eh[(3 * (size_t)(j) + 2) * es]
	// first row (ksw2_extz.c:32-35, ksw2_extd.c:33-41); kcalloc zero-fills eh, which only matters for e2 of ksw_extz (unused)
	EH(0) = 0; EE(0) = -gapoe - gapoe; EE2(0) = dual ? -gapoe2 - gapoe2 : 0;
	int j;
	for (j = 1; j <= qlen && j <= w; ++j) {
		if (!dual) { EH(j) = -(gapoe + gape * (j - 1)); EE(j) = -(gapoe + gapoe + gape * j); EE2(j) = 0; }
		else {
			const int a = -(gapo + gape * j), b = -(gapo2 + gape2 * j), tmp = -(gapoe + gape * j) > -(gapoe2 + gape2 * j) ? -(gapoe + gape * j) : -(gapoe2 + gape2 * j);
			EH(j) = a > b ? a : b; EE(j) = tmp - gapoe; EE2(j) = tmp - gapoe2;
		}
	}
	for (; j <= qlen; ++j) { EH(j) = EE(j) = KS_NEG_INF; EE2(j) = dual ? KS_NEG_INF : 0; }
	for (int i = 0; i < tlen; ++i) {
		int max = KS_NEG_INF;
		const int8_t *mrow = P.mat + (size_t)target[i] * P.m;
		const int st = i > w ? i - w : 0, en = i + w < qlen - 1 ? i + w : qlen - 1;
		int h1, f, f2 = 0;
		if (!dual) { h1 = st > 0 ? KS_NEG_INF : -(gapoe + gape * i); f = st > 0 ? KS_NEG_INF : -(gapoe + gapoe + gape * i); }
		else {
			const int tmp = -(gapoe + gape * i) > -(gapoe2 + gape2 * i) ? -(gapoe + gape * i) : -(gapoe2 + gape2 * i);
			h1 = st > 0 ? KS_NEG_INF : tmp; f = st > 0 ? KS_NEG_INF : tmp - gapoe; f2 = st > 0 ? KS_NEG_INF : tmp - gapoe2;
		}
		uint8_t *zi = cig ? z + (size_t)i * (size_t)n_col : (uint8_t*)0;
		for (j = st; j <= en; ++j) {
			int h = EH(j), e = EE(j), e2 = dual ? EE2(j) : 0, d;
			EH(j) = h1;
			h += mrow[query[j]];
			if (!right) {             // left-aligned gaps: ties prefer H, then E (ksw2_extz.c:71-74, ksw2_extd.c:92-99); score-only = same values
				d = h >= e ? 0 : 1; h = h >= e ? h : e;
				d = h >= f ? d : 2; h = h >= f ? h : f;
				if (dual) { d = h >= e2 ? d : 3; h = h >= e2 ? h : e2; d = h >= f2 ? d : 4; h = h >= f2 ? h : f2; }
			} else {                  // right-aligned (:98-101 / :132-139)
				d = h > e ? 0 : 1; h = h > e ? h : e;
				d = h > f ? d : 2; h = h > f ? h : f;
				if (dual) { d = h > e2 ? d : 3; h = h > e2 ? h : e2; d = h > f2 ? d : 4; h = h > f2 ? h : f2; }
			}
			h1 = h;
			if (right && !dual) { max_j = max >= h ? max_j : j; max = max >= h ? max : h; }   // ksw2_extz.c:103-104 (only this variant uses >=)
			else { max_j = max > h ? max_j : j; max = max > h ? max : h; }
			h -= gapoe; e -= gape;
			if (!right) { d |= e > h ? 0x08 : 0; e = e > h ? e : h; } else { d |= e >= h ? 0x08 : 0; e = e >= h ? e : h; }
			EE(j) = e;
			f -= gape;
			if (!right) { d |= f > h ? 0x10 : 0; f = f > h ? f : h; } else { d |= f >= h ? 0x10 : 0; f = f >= h ? f : h; }
			if (dual) {
				const int h2 = h1 - gapoe2;
				e2 -= gape2;
				if (!right) { d |= e2 > h2 ? 0x20 : 0; e2 = e2 > h2 ? e2 : h2; } else { d |= e2 >= h2 ? 0x20 : 0; e2 = e2 >= h2 ? e2 : h2; }
				EE2(j) = e2;
				f2 -= gape2;
				if (!right) { d |= f2 > h2 ? 0x40 : 0; f2 = f2 > h2 ? f2 : h2; } else { d |= f2 >= h2 ? 0x40 : 0; f2 = f2 >= h2 ? f2 : h2; }
			}
			if (cig) zi[j - st] = (uint8_t)d;
		}
		if (gg) j = en + 1;                                   // ksw_gg writes eh[en] with its exclusive en (ksw2_gg.c:92), always inside eh[]
		if (j <= qlen) { EH(j) = h1; EE(j) = KS_NEG_INF; }    // j == en + 1 (:113 / :151); e2 of that column is left as it is.  (A row whose band starts
		                                                      // beyond the query end makes the reference write past eh[]; not replicated.)
		if (en == qlen - 1 && EH(qlen) > ez.mqe) { ez.mqe = EH(qlen); ez.mqe_t = i; }
		if (i == tlen - 1) { ez.mte = max; ez.mte_q = max_j; }
		if (ks_rows_zdrop(ez, max, i, max_j, P.zdrop, dual ? gape2 : gape)) { ez.n_diag = i + 1; break; }
		if (i == tlen - 1 && en == qlen - 1) ez.score = EH(qlen);
	}
	if (gg) { const int sc = EH(qlen); ks_ez_reset(ez); ez.score = sc; ez.n_diag = tlen; }   // ksw2_gg.c:96: the score is all ksw_gg reports
#undef EH
#undef EE
#undef EE2
}

// where ksw_extz / ksw_extd start the traceback (ksw2_extz.c:127-132): no reach_end branch
KS_HD void ks_rows_pick_start(const KsRowsParams &P, int qlen, int tlen, const KsEz &ez, KsResult &o)
{
	o.tb_i = o.tb_j = -1; o.reach_end = 0;
	if (P.flag & KSF_SCORE_ONLY) return;
	if (!ez.zdropped && !(P.flag & KSF_EXTZ_ONLY)) { o.tb_i = tlen - 1; o.tb_j = qlen - 1; }
	else if (ez.max_t >= 0 && ez.max_q >= 0) { o.tb_i = ez.max_t; o.tb_j = ez.max_q; }
}

// ksw_backtrack, is_rot == 0, off_end == NULL, min_intron_len == 0 (ksw2.h:129-161).  out == 0: count only.
// A cell right of the band (j > i + w), which the reference would read from never-written heap, is treated as a forced
// deletion so that the walk is deterministic; a traceback that starts inside the band never gets there.
KS_HD int ks_rows_traceback(const KsRowsParams &P, int qlen, int tlen, const uint8_t *z, int i, int j, uint32_t *out, int n_total)
{
	const int w = ks_rows_w(P, qlen, tlen);
	const long long n_col = ks_rows_ncol(P, qlen, tlen);
	const bool rev = (P.flag & KSF_REV_CIGAR) != 0;
	int state = 0, n = 0, cur_op = -1, cur_len = 0;
#define KS_EMIT(OP, LEN) do { const int op_ = (OP), len_ = (LEN); \
		if (op_ == cur_op) cur_len += len_; \
		else { if (cur_op >= 0) { if (out) out[rev ? n : n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; } cur_op = op_; cur_len = len_; } } while (0)
	while (i >= 0 && j >= 0) {
		const int st = i > w ? i - w : 0, en = i + w < qlen - 1 ? i + w : qlen - 1;
		int force = -1;
		if (j < st) force = 2;
		else if (j > en) force = 1;
		const uint32_t d = force < 0 ? z[(size_t)i * (size_t)n_col + (size_t)(j - st)] : 0u;
		if (state == 0) state = d & 7;
		else if (!((d >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = d & 7;
		if (force >= 0) state = force;
		if (state == 0) { KS_EMIT(0, 1); --i; --j; }
		else if (state == 1 || state == 3) { KS_EMIT(2, 1); --i; }
		else { KS_EMIT(1, 1); --j; }
	}
	if (i >= 0) KS_EMIT(2, i + 1);
	if (j >= 0) KS_EMIT(1, j + 1);
	if (cur_op >= 0) { if (out) out[rev ? n : n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; }
#undef KS_EMIT
	return n;
}
