// ksw2_extf2.cuh -- ksw_extf2_sse (ksw2_extf2_sse.c:11-98): linear gap cost, u/v-only difference recurrence with
// X-drop on one tracked cell, score only.  GPU code path: one thread per pair, in-order over the anti-diagonals, lane by lane, with
// the pair's state in a global scratch area that has the reference's flat layout  u | v | s | sf (target copy) | qr (reversed
// query), all zero-initialised: the layout is observable (the unaligned 16-lane score chunks read past sf into qr and write past
// s into sf), see the oracle's kso_extf2 for the same restatement on the CPU.
#pragma once
#include "ksw2_pair.cuh"

struct KsExtfParams { int mch, mis, e, w, xdrop; };       // int8 values of the reference's arguments (mis as passed)

KS_HD size_t ks_extf2_scratch_bytes(int qlen, int tlen) { return (size_t)(((tlen + 15) / 16) * 4 + (qlen + 15) / 16 + 2) * 16; }

KS_HD void ks_extf2(const KsExtfParams &P, const uint8_t *query, int qlen, const uint8_t *target, int tlen, uint8_t *mem, KsEz &ez)
{
	const int tlen_ = (tlen + 15) / 16, L = tlen_ * 16;
	const int sc_mis = (int8_t)(P.mis < 0 ? P.mis : -P.mis), e2 = (int8_t)(P.e * 2), e = P.e;
	const int w = P.w < 0 ? (tlen > qlen ? tlen : qlen) : P.w;
	uint8_t *U = mem, *V = U + L, *S = V + L, *SF = S + L, *QR = SF + L;
	const size_t total = ks_extf2_scratch_bytes(qlen, tlen);
	for (size_t i = 0; i < total; ++i) mem[i] = 0;
	for (int t = 0; t < qlen; ++t) QR[t] = query[qlen - 1 - t];
	for (int t = 0; t < tlen; ++t) SF[t] = target[t];
	ks_ez_reset(ez);
	int last_st = -1, last_en = -1, last_t = 0, H0 = 0, r, nd = -1;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = ks_imax(ks_imax(0, r - qlen + 1), (r - w + 1) >> 1), en = ks_imin(ks_imin(tlen - 1, r), (r + w) >> 1);
		if (st > en) break;
		const int st0 = st, en0 = en;
		const uint8_t *qrr = QR + (qlen - 1 - r);
		st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
		uint8_t carry = (st > 0 && st - 1 >= last_st && st - 1 <= last_en) ? V[st - 1] : 0;
		if (en >= r) U[r] = 0;
		for (int t = st0; t <= en0; t += 16) {              // chunk: all 16 loads, then the store (:53-61)
			uint8_t tmp[16];
			for (int k = 0; k < 16; ++k) tmp[k] = (uint8_t)(SF[t + k] == qrr[t + k] ? P.mch : sc_mis);
			for (int k = 0; k < 16; ++k) S[t + k] = tmp[k];
		}
		for (int t = st; t <= en; ++t) {
			const int vt1 = (int8_t)carry, ut = (int8_t)U[t];
			int z = (int8_t)((int8_t)S[t] + e2);
			carry = V[t];
			z = z > vt1 ? z : vt1;                                          // signed (:72)
			z = (uint8_t)z > (uint8_t)ut ? z : ut;                          // unsigned (:77)
			U[t] = (uint8_t)(z - vt1); V[t] = (uint8_t)(z - ut);
		}
		if (r > 0) {
			if (last_t >= st0 && last_t <= en0 && last_t + 1 >= st0 && last_t + 1 <= en0) {
				const int d0 = (int)V[last_t] - e, d1 = (int)U[last_t + 1] - e;
				if (d0 > d1) H0 += d0; else { H0 += d1; ++last_t; }
			} else if (last_t >= st0 && last_t <= en0) H0 += (int)V[last_t] - e;
			else { ++last_t; H0 += (int)U[last_t] - e; }
			if (H0 > ez.max) { ez.max = H0; ez.max_t = last_t; ez.max_q = r - last_t; }
			else if (P.xdrop >= 0 && ez.max - H0 > P.xdrop) { nd = r + 1; break; }
		} else { H0 = (int)V[0] - e - e; last_t = 0; }
		last_st = st; last_en = en;
	}
	ez.n_diag = nd >= 0 ? nd : r;                          // diagonals evaluated (cell accounting)
	if (r == qlen + tlen - 1) ez.score = H0; else ez.zdropped = 1;
}
