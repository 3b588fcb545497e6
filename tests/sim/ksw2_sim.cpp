// ksw2_sim.cpp -- TEST INFRASTRUCTURE ONLY: host build of the device engine (ksw2_tile.cuh /
// ksw2_pair.cuh compiled as plain C++).  It lets the CPU-only test suite fuzz the exact code the GPU
// threads run (one thread = one pair, no inter-thread communication, so a sequential loop is a faithful
// functional simulation) against the oracle.  The product library never contains or calls this.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../ksw2_b200/csrc/ksw2_pair.cuh"
#include "../../ksw2_b200/csrc/ksw2_params.h"
#include "../../ksw2_b200/csrc/ksw2_scalar.cuh"
#include "../../ksw2_b200/csrc/ksw2_rows.cuh"
#include "../../ksw2_b200/csrc/ksw2_extf2.cuh"
#include "../../ksw2_b200/csrc/ksw2_gg2.cuh"

static void run_scalar(const KsParams &P, const KsPair &c, KsResult &res, std::vector<uint32_t> &cig)
{
	const bool with_cig = !(P.flag & KSF_SCORE_ONLY);
	const int prows = ks_prows(c.qlen, c.tlen, c.w);
	std::vector<int8_t> scr(ks_scalar_scratch_bytes(c.tlen));
	std::vector<ks_u4> p(with_cig ? (size_t)c.tlen_ * prows : 1);
	memset(p.data(), 0x5A, p.size() * sizeof(ks_u4));
	KsEz ez;
	ks_pair_scalar(P, c, ez, scr.data(), (uint8_t*)p.data(), prows);
	ks_store_result(ez, res);
	ks_pick_start(P, c, ez, res);
	cig.clear();
	if (with_cig && res.tb_i >= 0) {
		int n = ks_traceback(P, c, (const uint8_t*)p.data(), prows, res.tb_i, res.tb_j, 0, 0);
		cig.resize(n);
		ks_traceback(P, c, (const uint8_t*)p.data(), prows, res.tb_i, res.tb_j, cig.data(), n);
		res.n_cigar = n;
	}
}

// row-wise entry points (ksw_extz / ksw_extd): the device function, element stride 1
static void run_rows(const KsRowsParams &R, const uint8_t *query, int qlen, const uint8_t *target, int tlen, KsResult &res, std::vector<uint32_t> &cig)
{
	const bool with_cig = !(R.flag & KSF_SCORE_ONLY);
	std::vector<int32_t> eh(ks_rows_eh_words(qlen));
	std::vector<uint8_t> z(with_cig ? ks_rows_z_bytes(R, qlen, tlen) + 1 : 1);
	memset(eh.data(), 0xA5, eh.size() * 4); memset(z.data(), 0x5A, z.size());
	KsEz ez;
	ks_rows_fill(R, query, qlen, target, tlen, eh.data(), 1, z.data(), ez);
	ks_store_result(ez, res);
	ks_rows_pick_start(R, qlen, tlen, ez, res);
	cig.clear();
	if (with_cig && res.tb_i >= 0) {
		int n = ks_rows_traceback(R, qlen, tlen, z.data(), res.tb_i, res.tb_j, 0, 0);
		cig.resize(n);
		ks_rows_traceback(R, qlen, tlen, z.data(), res.tb_i, res.tb_j, cig.data(), n);
		res.n_cigar = n;
	}
}

template<int KIND, int CIG>
static void run_one(const KsParams &P, const KsPair &c, int C, KsResult &res, std::vector<uint32_t> &cig)
{
	const int SW = ks_save_words(P, KsSaveWords<KIND>::value);
	std::vector<ks_u4> save((size_t)c.tlen_ * SW);
	std::vector<ks_u4> bufA(C > 0 ? C + 1 : 1), bufB(C > 0 ? C + 1 : 1), best(C > 0 ? C : 1);
	const int prows = ks_prows(c.qlen, c.tlen, c.w);
	std::vector<ks_u4> p(KS_DIR(CIG) ? (size_t)c.tlen_ * prows : 1);
	memset(save.data(), 0xA5, save.size() * sizeof(ks_u4));      // poison: stale reads must not matter
	memset(p.data(), 0x5A, p.size() * sizeof(ks_u4));
	// one-off encoding of the pair (what ks_encode_kernel does on the device)
	std::vector<uint8_t> tenc((size_t)c.tlen_ * 16), qreg(ks_qenc_bytes(c.qlen));
	for (int i = 0; i < c.tlen_ * 16; ++i) tenc[(size_t)(i & ~15) + ks_perm_pos(i & 15)] = ks_enc_t(P, c.target, c.tlen, i);
	for (int i = -KS_QPADL; i < (int)qreg.size() - KS_QPADL; ++i) qreg[(size_t)(i + KS_QPADL)] = ks_enc_q(P, c.query, c.qlen, i);
	KsPair cc = c; cc.tenc = tenc.data(); cc.qenc = qreg.data() + KS_QPADL;
	KsEz ez;
	if (C < 0) {       // warp-cooperative driver, simulated lane by lane
		const int Cw = -C;
		if (Cw == 200000) {                                  // C = -200000: the ring schedule (banded pairs; the caller checks the band)
			std::vector<ks_u4> ring(256);
			memset(ring.data(), 0xC3, ring.size() * sizeof(ks_u4));
			KsWarpShared sh;
			ks_pair_fill_ring<KIND, CIG>(P, cc, &sh, save.data(), ring.data(), p.data(), prows);
			ez = sh.ez;
		} else {
		// panel heights above 100000 select the CTA-wide wavefront (64 lanes in the simulation): C = -(100000 + panel)
		const bool cta = Cw > 100000;
		const int Cp = cta ? Cw - 100000 : Cw;
		std::vector<ks_u4> ring(8 * 64), inw(66), wv(4 * (size_t)(Cp + 1));
		memset(ring.data(), 0xC3, ring.size() * sizeof(ks_u4)); memset(wv.data(), 0x3C, wv.size() * sizeof(ks_u4)); memset(inw.data(), 0x99, inw.size() * sizeof(ks_u4));
		KsWarpShared sh;
		if (cta) ks_pair_fill_warp<KIND, CIG, 64>(P, cc, &sh, Cp, save.data(), ring.data(), inw.data(), wv.data(), p.data(), prows);
		else ks_pair_fill_warp<KIND, CIG, 32>(P, cc, &sh, Cp, save.data(), ring.data(), inw.data(), wv.data(), p.data(), prows);
		ez = sh.ez;
		}
	} else
	ks_pair_fill<KIND, CIG>(P, cc, ez, C, save.data(), bufA.data(), best.data(), 1, p.data(), prows);
	ks_store_result(ez, res);
	ks_pick_start(P, c, ez, res);
	cig.clear();
	if (KS_DIR(CIG) && res.tb_i >= 0) {
		int n = ks_traceback(P, c, (const uint8_t*)p.data(), prows, res.tb_i, res.tb_j, 0, 0);
		cig.resize(n);
		ks_traceback(P, c, (const uint8_t*)p.data(), prows, res.tb_i, res.tb_j, cig.data(), n);
		res.n_cigar = n;
	}
}

extern "C" int64_t kssim_run(int kind, int m, const int8_t *mat, int q, int e, int q2, int e2, int w, int zdrop, int end_bonus,
                             int flag, int noncan, int junc_bonus, int64_t n, const uint8_t *qcat, const int64_t *qoff,
                             const uint8_t *tcat, const int64_t *toff, const uint8_t *jcat, int C, int force_smode,
                             int32_t *res /* n x 12 */, int64_t *cig_off, uint32_t *cig_buf, int64_t cig_cap)
{
	KsParams P;
	std::vector<int8_t> smat((size_t)(m > 0 ? m * m : 1));
	if (kind == 7 || kind == 8) {                        // ksw_gg2 / ksw_gg2_sse
		KsGg2Params G; G.sse = kind == 8; G.m = m; G.q = (int8_t)q; G.e = (int8_t)e; G.w = w; G.mat = mat;
		const bool with = !(flag & KSF_SCORE_ONLY);
		int64_t tot = 0;
		for (int64_t i = 0; i < n; ++i) {
			const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
			KsResult r; KsEz ez; ks_ez_reset(ez);
			std::vector<uint32_t> cig;
			if (ql > 0 && tl > 0) {
				std::vector<int8_t> scr(ks_gg2_scratch_bytes(tl), (int8_t)0x77);
				std::vector<uint8_t> dir(with ? ks_gg2_dir_bytes(G, ql, tl) : 1, 0x5A);
				ez.score = ks_gg2_fill(G, qcat + qoff[i], ql, tcat + toff[i], tl, scr.data(), with ? dir.data() : (uint8_t*)0);
				if (with) { const int nc = ks_gg2_traceback(G, ql, tl, dir.data(), 0, 0); cig.resize(nc); ks_gg2_traceback(G, ql, tl, dir.data(), cig.data(), nc); }
			}
			ks_store_result(ez, r);
			int32_t *o = res + i * 12;
			o[0] = r.max; o[1] = r.zdropped; o[2] = r.max_q; o[3] = r.max_t; o[4] = r.mqe; o[5] = r.mqe_t; o[6] = r.mte; o[7] = r.mte_q;
			o[8] = r.score; o[9] = (int32_t)cig.size(); o[10] = 0; o[11] = r.n_diag;
			if (cig_off) {
				cig_off[i] = tot;
				if (tot + (int64_t)cig.size() <= cig_cap && cig_buf) memcpy(cig_buf + tot, cig.data(), cig.size() * 4);
				tot += (int64_t)cig.size();
			}
		}
		if (cig_off) cig_off[n] = tot;
		return (cig_off && tot > cig_cap) ? -1 : 0;
	}
	if (kind == 5) {                                     // ksw_extf2_sse: q = mch, q2 = mis, zdrop = xdrop
		KsExtfParams F; F.mch = (int8_t)q; F.mis = (int8_t)q2; F.e = (int8_t)e; F.w = w; F.xdrop = zdrop;
		for (int64_t i = 0; i < n; ++i) {
			const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
			KsResult r; KsEz ez; ks_ez_reset(ez);
			if (ql > 0 && tl > 0) {
				std::vector<uint8_t> mem(ks_extf2_scratch_bytes(ql, tl), 0xEE);
				ks_extf2(F, qcat + qoff[i], ql, tcat + toff[i], tl, mem.data(), ez);
			}
			ks_store_result(ez, r); r.tb_i = r.tb_j = -1; r.reach_end = 0;
			int32_t *o = res + i * 12;
			o[0] = r.max; o[1] = r.zdropped; o[2] = r.max_q; o[3] = r.max_t; o[4] = r.mqe; o[5] = r.mqe_t; o[6] = r.mte; o[7] = r.mte_q;
			o[8] = r.score; o[9] = 0; o[10] = 0; o[11] = r.n_diag;
			if (cig_off) cig_off[i] = 0;
		}
		if (cig_off) cig_off[n] = 0;
		return 0;
	}
	if (kind == 3 || kind == 4 || kind == 6) {
		KsRowsParams R; R.kind = kind == 3 ? KS_ROWZ : kind == 6 ? KS_ROWG : KS_ROWD; R.m = m; R.gapo = (int8_t)q; R.gape = (int8_t)e; R.gapo2 = (int8_t)q2; R.gape2 = (int8_t)e2;
		R.w = w; R.zdrop = zdrop; R.flag = flag; R.mat = mat;
		int64_t tot = 0;
		std::vector<uint32_t> cig;
		for (int64_t i = 0; i < n; ++i) {
			const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
			KsResult r; KsEz ez; ks_ez_reset(ez); ks_store_result(ez, r); r.tb_i = r.tb_j = -1; r.reach_end = 0;
			cig.clear();
			if (ql > 0 && tl > 0) run_rows(R, qcat + qoff[i], ql, tcat + toff[i], tl, r, cig);
			int32_t *o = res + i * 12;
			o[0] = r.max; o[1] = r.zdropped; o[2] = r.max_q; o[3] = r.max_t; o[4] = r.mqe; o[5] = r.mqe_t; o[6] = r.mte; o[7] = r.mte_q;
			o[8] = r.score; o[9] = r.n_cigar; o[10] = r.reach_end; o[11] = r.n_diag;
			if (cig_off) {
				cig_off[i] = tot;
				if (tot + (int64_t)cig.size() <= cig_cap && cig_buf) memcpy(cig_buf + tot, cig.data(), cig.size() * 4);
				tot += (int64_t)cig.size();
			}
		}
		if (cig_off) cig_off[n] = tot;
		return (cig_off && tot > cig_cap) ? -1 : 0;
	}
	const int st = ks_prepare_params(P, kind, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, noncan, junc_bonus, smat.data(), force_smode & 1);
	P.treload = (force_smode >> 1) & 1;          // bit 1 of force_smode: saved blocks without the coded target word
	P.mat = smat.data();
	int64_t tot = 0;
	std::vector<uint32_t> cig;
	for (int64_t i = 0; i < n; ++i) {
		const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
		KsResult r; KsEz ez; ks_ez_reset(ez); ks_store_result(ez, r); r.tb_i = r.tb_j = -1; r.reach_end = 0;
		cig.clear();
		if (st == KS_PREP_OK && ql > 0 && tl > 0) {
			KsPair c; ks_make_pair(c, P, qcat + qoff[i], ql, tcat + toff[i], tl, jcat ? jcat + toff[i] : 0);
			const int cg = (flag & KSF_SCORE_ONLY) ? 0 : (flag & KSF_RIGHT) ? 2 : 1;
#define GO(K, G) run_one<K, G>(P, c, C, r, cig)
			if ((flag & KSF_APPROX_MAX) && C == 0) run_scalar(P, c, r, cig);     // panel 0: the in-order scalar path (kept as a second opinion)
			else if (flag & KSF_APPROX_MAX) {                    // the approximate-max kernel variants (CIG + 4)
				if (kind == KS_Z) { if (cg == 0) GO(KS_Z, 4); else if (cg == 1) GO(KS_Z, 5); else GO(KS_Z, 6); }
				else if (kind == KS_D) { if (cg == 0) GO(KS_D, 4); else if (cg == 1) GO(KS_D, 5); else GO(KS_D, 6); }
				else { if (cg == 0) GO(KS_S, 4); else if (cg == 1) GO(KS_S, 5); else GO(KS_S, 6); }
			}
			else if (kind == KS_Z) { if (cg == 0) GO(KS_Z, 0); else if (cg == 1) GO(KS_Z, 1); else GO(KS_Z, 2); }
			else if (kind == KS_D) { if (cg == 0) GO(KS_D, 0); else if (cg == 1) GO(KS_D, 1); else GO(KS_D, 2); }
			else { if (cg == 0) GO(KS_S, 0); else if (cg == 1) GO(KS_S, 1); else GO(KS_S, 2); }
#undef GO
		}
		int32_t *o = res + i * 12;
		o[0] = r.max; o[1] = r.zdropped; o[2] = r.max_q; o[3] = r.max_t; o[4] = r.mqe; o[5] = r.mqe_t; o[6] = r.mte; o[7] = r.mte_q;
		o[8] = r.score; o[9] = r.n_cigar; o[10] = r.reach_end; o[11] = r.n_diag;
		if (cig_off) {
			cig_off[i] = tot;
			if (tot + (int64_t)cig.size() <= cig_cap && cig_buf) memcpy(cig_buf + tot, cig.data(), cig.size() * 4);
			tot += (int64_t)cig.size();
		}
	}
	if (cig_off) cig_off[n] = tot;
	return (cig_off && tot > cig_cap) ? -1 : 0;
}
