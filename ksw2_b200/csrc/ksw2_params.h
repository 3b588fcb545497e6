// ksw2_params.h -- host-side preparation of the per-batch constants (KsParams) from the caller's
// arguments, i.e. the set-up section of the reference entry points:
//   ksw2_extz2_sse.c:56-82, ksw2_extd2_sse.c:75-105, ksw2_exts2_sse.c:72-96.
#pragma once
#include <string.h>
#include "ksw2_tile.cuh"

enum { KS_PREP_OK = 0, KS_PREP_EARLY_OUT = 1 /* reference returns with a reset ez */ };

static inline int ks_w8(int v) { return (int)(int8_t)(uint8_t)(unsigned)v; }

// smat (m*m, host) receives the s-contribution table used by smode 1; may be NULL when the caller knows smode == 0.
static inline int ks_prepare_params(KsParams &P, int kind, int m, const int8_t *mat, int q, int e, int q2, int e2,
                                    int w, int zdrop, int end_bonus, int flag, int noncan, int junc_bonus,
                                    int8_t *smat, int force_smode)
{
	memset(&P, 0, sizeof P);
	P.kind = kind; P.flag = flag; P.m = m;
	q = (int8_t)q; e = (int8_t)e; q2 = (int8_t)q2; e2 = (int8_t)e2; noncan = (int8_t)noncan; junc_bonus = (int8_t)junc_bonus;
	if (kind == KS_Z ? m <= 0 : m <= 1) return KS_PREP_EARLY_OUT;
	if (kind == KS_S && (q2 <= q + e || e == 0)) return KS_PREP_EARLY_OUT;   // e == 0: the reference divides by e (:93); rejected
	const int qe_pre = q + e;
	if (kind == KS_D && q2 + e2 < q + e) { int t; t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
	int max_sc = mat[0], min_sc = m * m > 1 ? mat[1] : mat[0];
	for (int t = 1; t < m * m; ++t) { max_sc = max_sc > mat[t] ? max_sc : mat[t]; min_sc = min_sc < mat[t] ? min_sc : mat[t]; }
	if (-min_sc > 2 * (q + e)) return KS_PREP_EARLY_OUT;
	P.q = q; P.e = e; P.q2 = q2; P.e2 = e2;
	P.h0sub = kind == KS_Z ? 2 * (q + e) : kind == KS_D ? qe_pre : q + e;
	P.qe_sub = kind == KS_Z ? q + e : 0;
	P.w = kind == KS_S ? -1 : w;
	P.zdrop = zdrop; P.end_bonus = end_bonus;
	P.zdrop_e = kind == KS_Z ? e : kind == KS_D ? e2 : 0;
	if (kind == KS_D) {
		P.long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
		if (q2 + e2 + P.long_thres * e2 > q + e + P.long_thres * e) ++P.long_thres;
		P.long_diff = ks_w8(P.long_thres * (e - e2) - (q2 - q) - e2);
		P.e_far = ks_w8(-e2);
	} else if (kind == KS_S) {
		P.long_thres = (q2 - q) / e - 1;
		if (q2 > q + e + P.long_thres * e) ++P.long_thres;
		P.long_diff = ks_w8(P.long_thres * e - (q2 - q));
		P.e_far = 0;
	}
	P.clamp = kind == KS_Z ? ks_w8(mat[0] + 2 * (q + e)) : mat[0];
	P.init_a = kind == KS_Z ? 0 : ks_w8(-q - e);
	P.init_b = kind == KS_D ? ks_w8(-q2 - e2) : ks_w8(-q2);
	P.sz_init = kind == KS_Z ? ks_w8(2 * (q + e)) : 0;
	P.noncan = noncan; P.junc_bonus = junc_bonus; P.semi = (flag & KSF_SPLICE_FLANK) ? -noncan / 2 : 0;
	P.wild = m - 1;
	P.gen_sc = (flag & KSF_GENERIC_SC) ? 1 : 0;
	const int sc_mch = mat[0], sc_mis = m * m > 1 ? mat[1] : mat[0];
	const int sc_N = mat[m * m - 1] == 0 ? ks_w8(-(kind == KS_D ? e2 : e)) : mat[m * m - 1];
	const int add = kind == KS_Z ? 2 * (q + e) : 0;
	// 3-class look-up (smode 0): classes 0 = equal, 1..3 = different, 4..7 = wildcard involved
	int lut[8], zero_idx = -1;
	lut[0] = ks_w8(sc_mch + add); lut[1] = lut[2] = lut[3] = ks_w8(sc_mis + add); lut[4] = lut[5] = lut[6] = lut[7] = ks_w8(sc_N + add);
	for (int i = 7; i >= 0; --i) if (lut[i] >= 0) zero_idx = i;
	P.smode = (P.gen_sc || m < 2 || m > 5 || zero_idx < 0 || force_smode == 1) ? 1 : 0;
	P.lut_lo = P.lut_hi = 0;
	for (int i = 0; i < 4; ++i) { P.lut_lo |= ((uint32_t)lut[i] & 0xffu) << (8 * i); P.lut_hi |= ((uint32_t)lut[4 + i] & 0xffu) << (8 * i); }
	P.tlow = zero_idx >= 0 ? (uint32_t)(8 | zero_idx) : 0u;
	if (P.smode == 1 && smat) {
		for (int a = 0; a < m; ++a)
			for (int b = 0; b < m; ++b) {
				int s = P.gen_sc ? mat[a * m + b] : (a == m - 1 || b == m - 1) ? sc_N : a == b ? sc_mch : sc_mis;
				smat[a * m + b] = (int8_t)ks_w8(s + add);
			}
	}
	return KS_PREP_OK;
}

static inline void ks_make_pair(KsPair &c, const KsParams &P, const uint8_t *query, int qlen, const uint8_t *target, int tlen, const uint8_t *junc)
{
	const int mx = qlen > tlen ? qlen : tlen;
	c.query = query; c.target = target; c.junc = junc; c.tenc = 0; c.qenc = 0; c.qlen = qlen; c.tlen = tlen;
	c.w = (P.w < 0 || P.w > mx) ? mx : P.w;
	c.ndiag = qlen + tlen - 1; c.tlen_ = (tlen + 15) / 16;
}
