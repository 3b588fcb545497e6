// ksw2_scalar.cuh -- in-order scalar GPU path of the approximate-max mode (KSW_EZ_APPROX_MAX / KSW_EZ_APPROX_DROP): round 1's implementation of
// ksw2_extz2_sse.c:270-286, one thread per pair, diagonal by diagonal, lane by lane, state in a per-pair global scratch area.  Since round 2 the
// mode runs on the tile engine (ks_apx_step in ksw2_tile.cuh, 60 x faster); this kernel is kept as an independent second opinion, selected with
// KSW2B_SCALAR_APPROX=1 and by the host simulator at panel 0.  Semantics follow SURVEY.md Appendix A (16-lane rounding, stale score row, carry
// rules); the direction bytes go to the same [block][row][16] layout the traceback kernel reads.
#pragma once
#include "ksw2_pair.cuh"

// bytes of scratch one pair needs (arrays of L = tlen_*16 int8 lanes: u v x y x2 y2/donor acceptor s)
KS_HD size_t ks_scalar_scratch_bytes(int tlen) { const size_t L = (size_t)((tlen + 15) / 16) * 16; return 8 * L + 16; }

KS_HD int ks_sc_w8(int v) { return (int)(int8_t)(uint8_t)(unsigned)v; }

// s-contribution of lane t on diagonal r (what the reference's score loop would write), incl. the zero padding
// of both sequences (sf tail / qr tail) and the spill of target reads into the reversed query (:84-86,98-99)
KS_HD int ks_sc_sval(const KsParams &P, const KsPair &c, int r, int t)
{
	const int L = c.tlen_ * 16;
	int a, b;
	if (t < L) a = t < c.tlen ? c.target[t] : 0;
	else { const int i = t - L; a = i < c.qlen ? c.query[c.qlen - 1 - i] : 0; }          // sf[] is followed by qr[]
	{ const int i = c.qlen - 1 - r + t; b = (i >= 0 && i < c.qlen) ? c.query[c.qlen - 1 - i] : 0; }
	if (P.smode) return P.mat[a * P.m + b];
	const int cls = (a == P.wild || b == P.wild) ? 4 : (a == b ? 0 : 1);
	const uint32_t w = cls < 4 ? P.lut_lo : P.lut_hi;
	return (int)(int8_t)((w >> (8 * (cls & 3))) & 0xffu);
}

KS_HD void ks_pair_scalar(const KsParams &P, const KsPair &c, KsEz &ez, int8_t *scr, uint8_t *pbase, int prows)
{
	const int L = c.tlen_ * 16, kind = P.kind;
	const bool cig = !(P.flag & KSF_SCORE_ONLY), right = (P.flag & KSF_RIGHT) != 0, approx = (P.flag & KSF_APPROX_MAX) != 0;
	int8_t *U = scr, *V = U + L, *X = V + L, *Y = X + L, *X2 = Y + L, *Y2 = X2 + L, *AC = Y2 + L, *S = AC + L;   // Y2 doubles as exts2's donor
	ks_ez_reset(ez); ez.n_diag = c.ndiag;
	for (int t = 0; t < L; ++t) {
		U[t] = V[t] = X[t] = Y[t] = (int8_t)P.init_a; X2[t] = Y2[t] = (int8_t)P.init_b; AC[t] = 0; S[t] = (int8_t)P.sz_init;
		if (kind == KS_S) {
			int don, acc; ks_splice(P, c, t, don, acc);
			if (t >= c.tlen) don = acc = (P.flag & (KSF_SPLICE_FOR | KSF_SPLICE_REV)) ? (int8_t)-P.noncan : 0;
			Y2[t] = (int8_t)don; AC[t] = (int8_t)acc;
		}
	}
	int last_st = -1, last_en = -1, H0 = 0, last_t = 0;
	for (int r = 0; r < c.ndiag; ++r) {
		int st0, en0;
		if (!ks_geo(c, r, st0, en0)) { ez.zdropped = 1; ez.n_diag = r; break; }
		const int st = st0 & ~15, en = en0 | 15;
		int cx, cv, cx2;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) { cx = X[st - 1]; cv = V[st - 1]; cx2 = X2[st - 1]; }
			else { cx = cv = P.init_a; cx2 = P.init_b; }
		} else { cx = P.init_a; cx2 = P.init_b; cv = ks_bnd(P, r); }
		if (en >= r) { Y[r] = (int8_t)P.init_a; if (kind == KS_D) Y2[r] = (int8_t)P.init_b; U[r] = (int8_t)ks_bnd(P, r); }
		{   // score row: chunks of 16 from st0 (overshooting en0), or exactly [st0,en0] with KSW_EZ_GENERIC_SC; writes past the array end are dropped
			const int wend = P.gen_sc ? en0 + 1 : st0 + 16 * ((en0 - st0) / 16 + 1);
			for (int t = st0; t < wend && t < L; ++t) S[t] = (int8_t)ks_sc_sval(P, c, r, t);
		}
		const bool sx = kind == KS_Z && (int8_t)cx < 0, sv = kind == KS_Z && (int8_t)cv < 0;   // sign-extending carry move (:146-147)
		for (int t = st; t <= en; ++t) {
			int xt = (int8_t)cx, vt = (int8_t)cv, x2t = (int8_t)cx2;
			if (t > st && t < st + 4) { if (sx) xt = -1; if (sv) vt = -1; }
			cx = X[t]; cv = V[t]; cx2 = X2[t];
			const int ut = U[t];
			int a = ks_sc_w8(xt + vt), b = ks_sc_w8(Y[t] + ut), z = S[t], d = 0;
			if (kind == KS_Z) {
				if (!right) { d = a > z ? 1 : 0; z = z > a ? z : a; if (b > z) d = 2; }
				else { d = z > a ? 0 : 1; z = z > a ? z : a; d = z > b ? d : 2; }
				z = (uint8_t)z > (uint8_t)b ? z : b;
				z = (uint8_t)z < (uint8_t)P.clamp ? z : P.clamp;
				z = (int8_t)z;
			} else {
				const int a2 = ks_sc_w8(x2t + vt), v3 = kind == KS_D ? a2 : ks_sc_w8(a2 + AC[t]), v4 = kind == KS_D ? ks_sc_w8(Y2[t] + ut) : 0;
				if (!right) {
					if (a > z) d = 1;
					z = z > a ? z : a;
					if (b > z) d = 2;
					z = z > b ? z : b;
					if (v3 > z) d = 3;
					z = z > v3 ? z : v3;
					if (kind == KS_D) { if (v4 > z) d = 4; z = z > v4 ? z : v4; }
				} else {
					d = z > a ? 0 : 1; z = z > a ? z : a;
					d = z > b ? d : 2; z = z > b ? z : b;
					d = z > v3 ? d : 3; z = z > v3 ? z : v3;
					if (kind == KS_D) { d = z > v4 ? d : 4; z = z > v4 ? z : v4; }
				}
				if (kind == KS_D) z = z < P.clamp ? z : P.clamp;
				const int a2p = ks_sc_w8(a2 - ks_sc_w8(z - P.q2));
				if (kind == KS_D) {
					const int b2p = ks_sc_w8(v4 - ks_sc_w8(z - P.q2));
					X2[t] = (int8_t)ks_sc_w8((a2p > 0 ? a2p : 0) - (P.q2 + P.e2)); Y2[t] = (int8_t)ks_sc_w8((b2p > 0 ? b2p : 0) - (P.q2 + P.e2));
					if (!right) { if (a2p > 0) d |= 0x20; if (b2p > 0) d |= 0x40; } else { if (!(0 > a2p)) d |= 0x20; if (!(0 > b2p)) d |= 0x40; }
				} else {
					const int don = Y2[t];
					X2[t] = (int8_t)ks_sc_w8((a2p > don ? a2p : don) - P.q2);
					if (!right) { if (a2p > don) d |= 0x20; } else { if (!(don > a2p)) d |= 0x20; }
				}
			}
			U[t] = (int8_t)ks_sc_w8(z - vt); V[t] = (int8_t)ks_sc_w8(z - ut);
			const int zq = ks_sc_w8(z - P.q), ap = ks_sc_w8(a - zq), bp = ks_sc_w8(b - zq);
			if (kind == KS_Z) { X[t] = (int8_t)(ap > 0 ? ap : 0); Y[t] = (int8_t)(bp > 0 ? bp : 0); }
			else { X[t] = (int8_t)ks_sc_w8((ap > 0 ? ap : 0) - (P.q + P.e)); Y[t] = (int8_t)ks_sc_w8((bp > 0 ? bp : 0) - (P.q + P.e)); }
			if (!right) { if (ap > 0) d |= 0x08; if (bp > 0) d |= 0x10; } else { if (!(0 > ap)) d |= 0x08; if (!(0 > bp)) d |= 0x10; }
			if (cig) {
				const int k = t >> 4;
				pbase[((size_t)k * prows + (size_t)(r - ks_rin(c, k))) * 16 + ks_perm_pos(t & 15)] = (uint8_t)d;
			}
		}
		// approximate max: follow one cell (ksw2_extz2_sse.c:270-286; extd2 :367-383; exts2 :385-401)
		(void)approx;
#define KS_UVS(ARR, i) (kind == KS_Z ? (int)(uint8_t)(ARR)[i] : (int)(ARR)[i])
		if (r > 0) {
			if (last_t >= st0 && last_t <= en0 && last_t + 1 >= st0 && last_t + 1 <= en0) {
				const int d0 = KS_UVS(V, last_t) - P.qe_sub, d1 = KS_UVS(U, last_t + 1) - P.qe_sub;
				if (d0 > d1) H0 += d0; else { H0 += d1; ++last_t; }
			} else if (last_t >= st0 && last_t <= en0) H0 += KS_UVS(V, last_t) - P.qe_sub;
			else { ++last_t; H0 += KS_UVS(U, last_t) - P.qe_sub; }
			if (kind == KS_Z && (P.flag & KSF_APPROX_DROP) && ks_zdrop(P, ez, H0, r, last_t)) { ez.n_diag = r + 1; break; }
		} else { H0 = KS_UVS(V, 0) - P.h0sub; last_t = 0; }
		if (kind != KS_Z && (P.flag & KSF_APPROX_DROP) && ks_zdrop(P, ez, H0, r, last_t)) { ez.n_diag = r + 1; break; }
#undef KS_UVS
		if (r == c.ndiag - 1 && en0 == c.tlen - 1) ez.score = H0;
		last_st = st; last_en = en;
	}
}
