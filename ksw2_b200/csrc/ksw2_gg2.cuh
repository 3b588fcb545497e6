// ksw2_gg2.cuh -- the two anti-diagonal GLOBAL alignment entry points of ksw2.h:89-90 as a GPU code path:
//   ksw_gg2      (ksw2_gg2.c:4-114)       scalar int8, exact band, signed compares, boundary tests by band geometry
//   ksw_gg2_sse  (ksw2_gg2_sse.c:11-126)  16-lane vectors (rounded band), unsigned max for b, no clamp, stale-neighbour test
// One thread per pair, in-order over the anti-diagonals (like ksw2_scalar.cuh); state in a per-pair global scratch (five int8
// arrays), direction bytes in ONE flat region with the reference's row pitch, zero-filled first (ksw_gg2 callocs it; ksw_gg2_sse
// mallocs it, so a traceback that leaves the band on the right is undefined there and deterministic here).
// Traceback = ksw_backtrack with is_rot == 1, off_end == NULL, not reversed (ksw2.h:129-161); off[r] is recomputed.
#pragma once
#include "ksw2_pair.cuh"

struct KsGg2Params { int sse, m, q, e, w; const int8_t *mat; };   // mat: device pointer, m*m

KS_HD int ks_gg2_w(const KsGg2Params &P, int qlen, int tlen) { return P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w; }
KS_HD size_t ks_gg2_pitch(const KsGg2Params &P, int qlen, int tlen)
{
	const int w = ks_gg2_w(P, qlen, tlen), n_col = w + 1 < tlen ? w + 1 : tlen;
	return P.sse ? (size_t)((n_col + 15) / 16 + 1) * 16 : (size_t)n_col;
}
KS_HD size_t ks_gg2_scratch_bytes(int tlen) { return (size_t)(((tlen + 15) / 16) * 16 + 16) * 5; }
KS_HD size_t ks_gg2_dir_bytes(const KsGg2Params &P, int qlen, int tlen) { return (size_t)(qlen + tlen) * ks_gg2_pitch(P, qlen, tlen) + 16; }
KS_HD int ks_gg2_w8(int v) { return (int)(int8_t)(uint8_t)(unsigned)v; }
KS_HD void ks_gg2_geo(int qlen, int tlen, int w, int r, int &st0, int &en0)
{
	st0 = ks_imax(ks_imax(0, r - qlen + 1), (r - w + 1) >> 1);
	en0 = ks_imin(ks_imin(tlen - 1, r), (r + w) >> 1);
}

// returns the score (H0 of the tracked last-row cell); dir == 0: score only
KS_HD int ks_gg2_fill(const KsGg2Params &P, const uint8_t *query, int qlen, const uint8_t *target, int tlen, int8_t *scr, uint8_t *dir)
{
	const int sse = P.sse, q = P.q, qe = P.q + P.e, qe2 = ks_gg2_w8(qe * 2), w = ks_gg2_w(P, qlen, tlen);
	const int L = ((tlen + 15) / 16) * 16 + 16;
	const size_t pitch = ks_gg2_pitch(P, qlen, tlen);
	int8_t *U = scr, *V = U + L, *X = V + L, *Y = X + L, *S = Y + L;
	for (int t = 0; t < 5 * L; ++t) scr[t] = 0;
	if (dir) { const size_t nb = ks_gg2_dir_bytes(P, qlen, tlen); for (size_t i = 0; i < nb; ++i) dir[i] = 0; }
	int H0 = 0, last_t = 0, last_st = -1, last_en = -1;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0, en0;
		ks_gg2_geo(qlen, tlen, w, r, st0, en0);
		int st = st0, en = en0, x1, v1;
		if (sse) { st = st0 / 16 * 16; en = (en0 + 16) / 16 * 16 - 1; }
		if (!sse) {                                                            // ksw2_gg2.c:36-43
			if (st != 0) { if (r > st + st + w - 1) x1 = v1 = 0; else { x1 = X[st - 1]; v1 = V[st - 1]; } }
			else { x1 = 0; v1 = r ? q : 0; }
			if (en != r) { if (r < en + en - w - 1) Y[en] = U[en] = 0; }
			else { Y[r] = 0; U[r] = (int8_t)(r ? q : 0); }
		} else {                                                               // ksw2_gg2_sse.c:54-60
			if (st > 0) { if (st - 1 >= last_st && st - 1 <= last_en) { x1 = X[st - 1]; v1 = V[st - 1]; } else x1 = v1 = 0; }
			else { x1 = 0; v1 = r ? q : 0; }
			if (en >= r) { Y[r] = 0; U[r] = (int8_t)(r ? q : 0); }
		}
		x1 = (int8_t)x1; v1 = (int8_t)v1;
		for (int t = st0; t <= en0; ++t) S[t] = P.mat[target[t] * P.m + query[r - t]];   // qr[t + qlen - 1 - r] = query[r - t]
		for (int t = st; t <= en; ++t) {
			int z = ks_gg2_w8(S[t] + qe2), a = ks_gg2_w8(x1 + v1), b = ks_gg2_w8(Y[t] + U[t]);
			int d = a > z ? 1 : 0;
			z = z > a ? z : a;
			d = b > z ? 2 : d;
			z = sse ? ((uint8_t)z > (uint8_t)b ? z : b) : (z > b ? z : b);
			const int u1 = U[t];
			U[t] = (int8_t)ks_gg2_w8(z - v1); v1 = V[t]; V[t] = (int8_t)ks_gg2_w8(z - u1);
			z = ks_gg2_w8(z - q); a = ks_gg2_w8(a - z); b = ks_gg2_w8(b - z);
			x1 = X[t];
			if (a > 0) d |= 0x08;
			X[t] = (int8_t)(a > 0 ? a : 0);
			if (b > 0) d |= 0x10;
			Y[t] = (int8_t)(b > 0 ? b : 0);
			if (dir) dir[(size_t)r * pitch + (size_t)(t - st)] = (uint8_t)d;
		}
		if (r > 0) {
			if (last_t >= st0 && last_t <= en0) H0 += (sse ? (int)(uint8_t)V[last_t] : (int)V[last_t]) - qe;
			else { ++last_t; H0 += (sse ? (int)(uint8_t)U[last_t] : (int)U[last_t]) - qe; }
		} else { H0 = (sse ? (int)(uint8_t)V[0] : (int)V[0]) - 2 * qe; last_t = 0; }
		last_st = st; last_en = en;
	}
	return H0;
}

// out == 0: count only.  Ops are produced from the end of the alignment and stored reversed (the reference reverses at the end).
KS_HD int ks_gg2_traceback(const KsGg2Params &P, int qlen, int tlen, const uint8_t *dir, uint32_t *out, int n_total)
{
	const int w = ks_gg2_w(P, qlen, tlen);
	const size_t pitch = ks_gg2_pitch(P, qlen, tlen), nb = ks_gg2_dir_bytes(P, qlen, tlen);
	int i = tlen - 1, j = qlen - 1, state = 0, n = 0, cur_op = -1, cur_len = 0;
#define KS_EMIT(OP, LEN) do { const int op_ = (OP), len_ = (LEN); \
		if (op_ == cur_op) cur_len += len_; \
		else { if (cur_op >= 0) { if (out) out[n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; } cur_op = op_; cur_len = len_; } } while (0)
	while (i >= 0 && j >= 0) {
		const int r = i + j;
		int st0, en0, force = -1;
		ks_gg2_geo(qlen, tlen, w, r, st0, en0);
		const int off = P.sse ? st0 / 16 * 16 : st0;
		if (i < off) force = 2;
		const size_t idx = (size_t)r * pitch + (size_t)(i - off);
		const uint32_t d = force < 0 ? (idx < nb ? dir[idx] : 0u) : 0u;
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
	if (cur_op >= 0) { if (out) out[n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; }
#undef KS_EMIT
	return n;
}
