// ksw2_pair.cuh -- per-alignment drivers on top of the tile engine: the panel sweep on ONE THREAD (ks_pair_fill), the diagonal-skewed wavefront
// on ONE WARP or one CTA in waves of NL blocks (ks_pair_fill_warp) and on the ring schedule for banded pairs (ks_pair_fill_ring); result
// selection, traceback.  See DESIGN.md section 3.
#pragma once
#include "ksw2_tile.cuh"

struct KsResult {             // device-side result record, 64 bytes
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar;
	int32_t tb_i, tb_j;       // traceback start cell (-1: no CIGAR)
	int32_t n_diag;           // anti-diagonals evaluated (reference semantics: up to and including the Z-drop diagonal)
	int64_t cigar_off;        // word offset of this pair's CIGAR in the batch CIGAR buffer
};

KS_HD void ks_ez_reset(KsEz &ez)   // ksw2.h:184-189
{
	ez.max = 0; ez.max_t = ez.max_q = ez.mqe_t = ez.mte_q = -1;
	ez.mqe = ez.mte = ez.score = KS_NEG_INF; ez.zdropped = 0; ez.n_diag = 0;
	ez.apx_H0 = 0; ez.apx_t = 0; ez.apx_r = 0;
}

// Fill: sweeps panels of C diagonals; inside a panel, blocks left to right.
//  save  : per-thread state area, SW 16-byte words per block, block k at save + k*SW*sstride (sstride in words between consecutive words)
//  cs: carry stream, C+1 entries (in place);  best: C entries; both with element stride sst (in 16-byte words)
//  pbase : direction bytes of this pair, [block][row][16]; prows rows per block (CIG != 0)
template<int KIND, int CIG>
KS_HD void ks_pair_fill(const KsParams &P, const KsPair &c, KsEz &ez, int C,
                        ks_u4 *save, ks_u4 *cs, ks_u4 *best, int sst, ks_u4 *pbase, int prows)
{
	const int SW = ks_save_words(P, KsSaveWords<KIND>::value);
	bool done = false;
	ks_ez_reset(ez);
	ez.n_diag = c.ndiag;
	for (int R = 0; R < c.ndiag && !done; R += C) {
		int Rend = ks_imin(R + C, c.ndiag), stop = -1, st0, en0;
		for (int r = R; r < Rend; ++r) if (!ks_geo(c, r, st0, en0)) { stop = r; break; }
		if (stop >= 0) Rend = stop;
		if (Rend > R) {
			ks_geo(c, R, st0, en0);        const int kmin = st0 >> 4;
			ks_geo(c, Rend - 1, st0, en0); const int kmax = en0 >> 4;
			if (R > 0 && kmin > 0) cs[0] = save[(size_t)(kmin - 1) * SW];
			for (int k = kmin; k <= kmax && !done; ++k) {
				const int ra = ks_imax(R, ks_rin(c, k)), rb = ks_imin(Rend - 1, ks_rout(c, k));
				if (ra > rb) continue;
				ks_tile<KIND, CIG>(P, c, ez, k, ra, rb, R, save + (size_t)k * SW, k > 0 ? save + (size_t)(k - 1) * SW : save, cs, best, sst,
				                   KS_DIR(CIG) ? pbase + (size_t)k * prows : (ks_u4*)0, done);
			}
		}
		if (stop >= 0 && !done) { ez.zdropped = 1; ez.n_diag = stop; done = true; }     // band narrower than |tlen-qlen| (:111-114)
	}
}

// Where the traceback starts (ksw2_extz2_sse.c:292-303, exts2 :409-412)
KS_HD void ks_pick_start(const KsParams &P, const KsPair &c, const KsEz &ez, KsResult &o)
{
	o.tb_i = o.tb_j = -1; o.reach_end = 0;
	if (P.flag & KSF_SCORE_ONLY) return;
	if (!ez.zdropped && !(P.flag & KSF_EXTZ_ONLY)) { o.tb_i = c.tlen - 1; o.tb_j = c.qlen - 1; }
	else if (P.kind != KS_S && !ez.zdropped && (P.flag & KSF_EXTZ_ONLY) && ez.mqe + P.end_bonus > ez.max) {
		o.reach_end = 1; o.tb_i = ez.mqe_t; o.tb_j = c.qlen - 1;
	} else if (ez.max_t >= 0 && ez.max_q >= 0) { o.tb_i = ez.max_t; o.tb_j = ez.max_q; }
}

KS_HD void ks_store_result(const KsEz &ez, KsResult &o)
{
	o.max = ez.max; o.zdropped = ez.zdropped; o.max_q = ez.max_q; o.max_t = ez.max_t; o.mqe = ez.mqe; o.mqe_t = ez.mqe_t;
	o.mte = ez.mte; o.mte_q = ez.mte_q; o.score = ez.score; o.n_cigar = 0; o.cigar_off = 0; o.n_diag = ez.n_diag;
}

// direction byte of cell (r, t): [block][row][16], byte order inside a row = lanes 0,8,1,9 | 2,10,3,11 | 4,12,5,13 | 6,14,7,15
KS_HD uint32_t ks_dir(const KsPair &c, const uint8_t *pbase, int prows, int r, int t)
{
	const int k = t >> 4, L = t & 15, i = L & 7;
	const size_t row = (size_t)k * prows + (size_t)(r - ks_rin(c, k));
	return pbase[row * 16 + (size_t)((i >> 1) * 4 + (i & 1) * 2 + (L >> 3))];
}

// Traceback state machine (ksw2.h:129-161, is_rot branch).  out == 0: only count the run-length ops.
// Ops are produced from the alignment end towards its start; the reference reverses them unless KSW_EZ_REV_CIGAR.
KS_HD int ks_traceback(const KsParams &P, const KsPair &c, const uint8_t *pbase, int prows, int i, int j, uint32_t *out, int n_total)
{
	const int min_intron = P.kind == KS_S ? P.long_thres : 0;
	const bool rev = (P.flag & KSF_REV_CIGAR) != 0;
	// KSW_EZ_EQX (extd2 only, ksw2_extd2_sse.c:399-406): M becomes '=' / 'X'; positions count from the first op of the array
	// (ksw_cigar2eqx, ksw2.h:163-182 -- intended semantics, the reference's own post-pass is broken in this commit)
	const bool eqx = P.kind == KS_D && (P.flag & KSF_EQX) && c.query && c.target;
	const int i0 = i, j0 = j;
	int state = 0, n = 0, cur_op = -1, cur_len = 0;
#define KS_EMIT(OP, LEN) do { const int op_ = (OP), len_ = (LEN); \
		if (op_ == cur_op) cur_len += len_; \
		else { if (cur_op >= 0) { if (out) out[rev ? n : n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; } cur_op = op_; cur_len = len_; } } while (0)
	while (i >= 0 && j >= 0) {
		const int r = i + j;
		int st0, en0, force = -1;
		uint32_t d;
		ks_geo(c, r, st0, en0);
		if (i < (st0 & ~15)) force = 2;
		if (i > (en0 | 15)) force = 1;
		d = force < 0 ? ks_dir(c, pbase, prows, r, i) : 0u;
		if (state == 0) state = d & 7;
		else if (!((d >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = d & 7;
		if (force >= 0) state = force;
		if (state == 0) {
			int op = 0;
			if (eqx) op = (rev ? c.target[i0 - i] == c.query[j0 - j] : c.target[i] == c.query[j]) ? 7 : 8;
			KS_EMIT(op, 1); --i; --j;
		}
		else if (state == 1 || (state == 3 && min_intron <= 0)) { KS_EMIT(2, 1); --i; }
		else if (state == 3 && min_intron > 0) { KS_EMIT(3, 1); --i; }
		else { KS_EMIT(1, 1); --j; }
	}
	if (i >= 0) KS_EMIT(min_intron > 0 && i >= min_intron ? 3 : 2, i + 1);
	if (j >= 0) KS_EMIT(1, j + 1);
	if (cur_op >= 0) { if (out) out[rev ? n : n_total - 1 - n] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; ++n; }
#undef KS_EMIT
	return n;
}

// ---------------------------------------------------------------------------------------------------------
// Warp-cooperative fill: ONE WARP per alignment.  Inside a panel of C diagonals the blocks kb..kb+31 of a "wave" sit on the
// 32 lanes; lane l evaluates diagonal r = R + tau - l at time step tau, i.e. one diagonal behind its left neighbour, whose
// carry / arg-max records it picks up from a 4-deep ring in shared memory.  Lane 31 streams its records to the next wave
// (blocks kb+32..), lane 0 reads the previous wave's.  One __syncwarp per time step; no atomics, no inter-warp traffic.
// Same tiles, same order constraints, same results as the thread-per-alignment sweep -- used when a batch has too few
// (long) pairs to fill the GPU with one thread each.
//   ring : 32 lanes x 4 slots x {carry, best}            (ks_u4[256], shared memory, per warp)
//   inw  : window of 32 records of the previous wave's stream (ks_u4[64], shared memory, per warp)
//   wv   : 2 x (C+1) x {carry, best} inter-wave streams   (ks_u4[4*(C+1)], GLOBAL memory, per warp): lane 31 writes one record per step
//          (fire and forget), the whole warp refills lane 0's window every 32 steps with one coalesced load -- so the panel height C is
//          not bounded by shared memory and the skew of the wavefront (31 idle steps per wave and panel) is amortised over a tall panel
//   ezs  : the ksw_extz_t scalars + stop flag, shared by the lanes (per warp)
// NL = 32: one warp per alignment (4 alignments per CTA).  NL = 64 .. 256: the same wavefront over ALL the threads of a CTA (one alignment per
// CTA, __syncthreads per step): for the lone long pair, whose latency is what counts (BASELINE config 1; a single 16.5 kb pair on one warp
// takes longer than on one CPU core).
struct KsWarpShared { KsEz ez; int done; int pad[2]; };

#if defined(__CUDA_ARCH__)
#define KS_LDCG(p) __ldcg(p)                 // the inter-wave streams are written and read by different lanes of the warp: read them at L2
#define KS_SYNCWARP() do { if (NL == 32) __syncwarp(); else __syncthreads(); } while (0)
#define KS_LANE_LOOP(l) const int l = NL == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
#define KS_LANE_END
#else
#define KS_LDCG(p) (*(p))
#define KS_SYNCWARP()
#define KS_LANE_LOOP(l) for (int l = 0; l < NL; ++l) {
#define KS_LANE_END }
#endif

template<int KIND, int CIG, int NL>
#if defined(__CUDA_ARCH__)
__device__ __forceinline__
#else
static inline
#endif
void ks_pair_fill_warp(const KsParams &P, const KsPair &c, KsWarpShared *ezs, int C, ks_u4 *save, ks_u4 *ring, ks_u4 *inw, ks_u4 *wv, ks_u4 *pbase, int prows)
{
	const int SW = ks_save_words(P, KsSaveWords<KIND>::value);
#if defined(__CUDA_ARCH__)
	KsTile<KIND> T;                       // this lane's tile
#define KS_T(l) T
	bool act = false;
#define KS_ACT(l) act
	int fa = 0, fb = -1;                  // this lane's interior diagonals (ks_fast_range)
#define KS_FA(l) fa
#define KS_FB(l) fb
#else
	static thread_local KsTile<KIND> Ts[NL];   // host simulation: one tile context per simulated lane
#define KS_T(l) Ts[l]
	bool acts[NL];
#define KS_ACT(l) acts[l]
	int fas[NL], fbs[NL];
#define KS_FA(l) fas[l]
#define KS_FB(l) fbs[l]
#endif
	{ KS_LANE_LOOP(l) if (l == 0) { ks_ez_reset(ezs->ez); ezs->ez.n_diag = c.ndiag; ezs->done = 0; } KS_LANE_END }
	KS_SYNCWARP();
	ks_u4 *win = wv, *wout = wv + 2 * (size_t)(C + 1);
	for (int R = 0; R < c.ndiag && !ezs->done; R += C) {
		int Rend = ks_imin(R + C, c.ndiag), stop = -1, st0, en0;
		for (int r = R; r < Rend; ++r) if (!ks_geo(c, r, st0, en0)) { stop = r; break; }
		if (stop >= 0) Rend = stop;
		if (Rend > R) {
			ks_geo(c, R, st0, en0);        const int kmin = st0 >> 4;
			ks_geo(c, Rend - 1, st0, en0); const int kmax = en0 >> 4;
			{ KS_LANE_LOOP(l) if (l == 0 && R > 0 && kmin > 0) win[0] = save[(size_t)(kmin - 1) * SW]; KS_LANE_END }
			for (int kb = kmin; kb <= kmax && !ezs->done; kb += NL) {
				{ KS_LANE_LOOP(l)
					const int k = kb + l;
					KS_ACT(l) = false; KS_FA(l) = 0; KS_FB(l) = -1;
					if (k <= kmax) {
						const int ra = ks_imax(R, ks_rin(c, k)), rb = ks_imin(Rend - 1, ks_rout(c, k));
						if (ra <= rb) {
							ks_u4 seed;
							ks_tile_begin<KIND>(P, c, KS_T(l), k, ra, rb, save + (size_t)k * SW, seed);
							ks_fast_range(c, k, ra, rb, KS_FA(l), KS_FB(l));
							ring[(l * 4 + ((R - 1) & 3)) * 2] = seed;
							if (l == NL - 1) wout[0] = seed;
							KS_ACT(l) = true;
						}
					}
				KS_LANE_END }
				KS_SYNCWARP();
				const int nstep = (Rend - R) + NL - 1, nrec = Rend - R;          // records 0 .. nrec of the incoming stream
				for (int tau = 0; tau < nstep; ++tau) {
					// Lane 0 reads record tau+1 of the previous wave's stream at step tau (and record tau as "previous diagonal").  Every 32 steps the
					// warp refills the window inw[0..31] = records tau+1 .. tau+32 with one coalesced load; slot 32 keeps record tau across the refill.
					if ((tau & 31) == 0) {
						{ KS_LANE_LOOP(l) if (l == 0) inw[64] = tau ? inw[62] : KS_LDCG(win); KS_LANE_END }
						KS_SYNCWARP();
						{ KS_LANE_LOOP(l) const int idx = tau + 1 + l; if (l < 32 && idx <= nrec) { inw[l * 2] = KS_LDCG(win + (size_t)idx * 2); inw[l * 2 + 1] = KS_LDCG(win + (size_t)idx * 2 + 1); } KS_LANE_END }
						KS_SYNCWARP();
					}
					// If every lane that has a diagonal to do this step is strictly inside the band, the whole warp takes the interior fast
					// step (warp-uniform choice: no divergence).  With bands many blocks wide that is nearly every step of nearly every wave.
					// approximate-max mode: every lane works from the same snapshot of the tracker (one lane at most advances it in a step)
					const int axH0 = ezs->ez.apx_H0, axT = ezs->ez.apx_t, axR = ezs->ez.apx_r;
					if (KS_APX(CIG)) KS_SYNCWARP();
					bool allfast = true;
#if defined(__CUDA_ARCH__)
					{ const int r = R + tau - (NL == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x);
					  const bool busy = act && r >= T.ra && r <= T.rb, okf = !busy || (r >= fa && r <= fb);
					  allfast = NL == 32 ? __all_sync(0xffffffffu, okf) != 0 : __syncthreads_and(okf) != 0; }
#else
					for (int l = 0; l < NL; ++l) { const int r = R + tau - l; if (KS_ACT(l) && r >= KS_T(l).ra && r <= KS_T(l).rb && !(r >= KS_FA(l) && r <= KS_FB(l))) allfast = false; }
#endif
					{ KS_LANE_LOOP(l)
						const int r = R + tau - l;
						if (KS_ACT(l) && r >= KS_T(l).ra && r <= KS_T(l).rb && !ezs->done) {
							const int k = kb + l;
							ks_u4 cprev, ccur, bin, co, bo;
							if (l == 0) { cprev = (tau & 31) ? inw[((tau - 1) & 31) * 2] : inw[64]; ccur = inw[(tau & 31) * 2]; bin = inw[(tau & 31) * 2 + 1]; }
							else { const ks_u4 *lr = ring + (size_t)(l - 1) * 8; cprev = lr[((r - 1) & 3) * 2]; ccur = lr[(r & 3) * 2]; bin = lr[(r & 3) * 2 + 1]; }
							bool zs = false;
							if (allfast) {
								const int st0 = ks_imax(ks_imax(0, r - c.qlen + 1), (r - c.w + 1) >> 1);
								zs = ks_tile_step_fast<KIND, CIG>(P, c, ezs->ez, KS_T(l), r, st0, cprev, ccur.x, bin, co, bo, KS_DIR(CIG) ? pbase + (size_t)k * prows : (ks_u4*)0, axH0, axT, axR);
							} else
							zs = ks_tile_step_mixed<KIND, CIG>(P, c, ezs->ez, KS_T(l), r, r >= KS_FA(l) && r <= KS_FB(l), cprev, ccur, bin, k > 0 ? save + (size_t)(k - 1) * SW : save, co, bo,
							                                   KS_DIR(CIG) ? pbase + (size_t)k * prows : (ks_u4*)0, axH0, axT, axR);
							ring[(l * 4 + (r & 3)) * 2] = co; ring[(l * 4 + (r & 3)) * 2 + 1] = bo;
							if (l == NL - 1) { wout[(size_t)(r - R + 1) * 2] = co; wout[(size_t)(r - R + 1) * 2 + 1] = bo; }
							if (r == KS_T(l).rb) save[(size_t)k * SW] = co;      // persist the last carry at once: the block on the right may need it this panel
							if (zs) ezs->done = 1;
						}
					KS_LANE_END }
					KS_SYNCWARP();
				}
				{ KS_LANE_LOOP(l) if (KS_ACT(l) && !ezs->done) ks_tile_end<KIND>(P, c, KS_T(l), save + (size_t)(kb + l) * SW); KS_LANE_END }
				KS_SYNCWARP();
				{ ks_u4 *t = win; win = wout; wout = t; }
			}
		}
		{ KS_LANE_LOOP(l) if (l == 0 && stop >= 0 && !ezs->done) { ezs->ez.zdropped = 1; ezs->ez.n_diag = stop; ezs->done = 1; } KS_LANE_END }
		KS_SYNCWARP();
	}
#undef KS_T
#undef KS_ACT
#undef KS_FA
#undef KS_FB
}

// ---------------------------------------------------------------------------------------------------------
// Ring schedule of the warp-cooperative fill, for BANDED pairs (effective band <= KS_RING_MAX_W, i.e. at most ~33 blocks of a diagonal in
// the band): block k lives on lane k & 31 for its WHOLE life (diagonals rin(k) .. rout(k)) and evaluates diagonal r at time step
// tau = r + k.  A lane is done with block k before block k + 32 enters the band (rin(k+32) + 32 > rout(k) whenever 2w < 1026), so a lane
// holds one block at a time, ~31 of the 32 lanes are busy at every step, and there are no panels: no state is ever saved or restored and
// no inter-wave stream exists -- the ring of records simply wraps from lane 31 to lane 0.  (The wave schedule above keeps only ~40 % of
// its lanes busy on such bands: 33 blocks need two waves of 32.)  Diagonals are finalised in ascending order (the block that holds en0
// moves right as r grows), so Z-drop and the approximate tracker behave as in the other drivers.  Every step is the mixed step: the
// band's two edge blocks are always among the lanes.
#define KS_RING_MAX_W 512
template<int KIND, int CIG>
#if defined(__CUDA_ARCH__)
__device__ __forceinline__
#else
static inline
#endif
void ks_pair_fill_ring(const KsParams &P, const KsPair &c, KsWarpShared *ezs, ks_u4 *save, ks_u4 *ring, ks_u4 *pbase, int prows)
{
	const int NL = 32;
	const int SW = ks_save_words(P, KsSaveWords<KIND>::value);
#if defined(__CUDA_ARCH__)
	KsTile<KIND> T;
#define KS_T(l) T
	int kcur = 0, fa = 0, fb = -1;
	bool act = false;
#define KS_K(l) kcur
#define KS_ACT(l) act
#define KS_FA(l) fa
#define KS_FB(l) fb
#else
	static thread_local KsTile<KIND> Ts[32];
	int kcurs[32], fas[32], fbs[32]; bool acts[32];
#define KS_T(l) Ts[l]
#define KS_K(l) kcurs[l]
#define KS_ACT(l) acts[l]
#define KS_FA(l) fas[l]
#define KS_FB(l) fbs[l]
#endif
	{ KS_LANE_LOOP(l) if (l == 0) { ks_ez_reset(ezs->ez); ezs->ez.n_diag = c.ndiag; ezs->done = 0; } KS_LANE_END }
	// diagonals before the band runs empty (band narrower than |tlen - qlen|, :111-114)
	int nd = c.ndiag;
	{ int st0, en0; for (int r = 0; r < c.ndiag; ++r) if (!ks_geo(c, r, st0, en0)) { nd = r; break; } }
	int kmax = 0;
	if (nd > 0) { int st0, en0; ks_geo(c, nd - 1, st0, en0); kmax = en0 >> 4; }
	{ KS_LANE_LOOP(l) KS_K(l) = l - NL; KS_ACT(l) = false; KS_FA(l) = 0; KS_FB(l) = -1; KS_LANE_END }    // (the first step moves every lane to its block l)
	const int tau_end = nd > 0 ? (nd - 1) + kmax : -1;
	for (int tau = 0; tau <= tau_end && !ezs->done; ++tau) {
		const int axH0 = ezs->ez.apx_H0, axT = ezs->ez.apx_t, axR = ezs->ez.apx_r;
		if (KS_APX(CIG)) KS_SYNCWARP();
		{ KS_LANE_LOOP(l)
			// move on to the lane's next block (k + 32, k + 64, ...) once the current one has left the band
			if (!KS_ACT(l) || tau - KS_K(l) > KS_T(l).rb) {
				KS_ACT(l) = false;
				while (KS_K(l) + NL <= kmax) {
					const int k = (KS_K(l) += NL);
					const int ra = ks_rin(c, k), rb = ks_imin(nd - 1, ks_rout(c, k));
					if (ra <= rb) { ks_u4 seed; ks_tile_begin<KIND>(P, c, KS_T(l), k, ra, rb, save + (size_t)k * SW, seed); ks_fast_range(c, k, ra, rb, KS_FA(l), KS_FB(l)); KS_ACT(l) = true; break; }
				}
			}
			const int k = KS_K(l), r = tau - k;
			if (KS_ACT(l) && r >= KS_T(l).ra && r <= KS_T(l).rb && !ezs->done) {
				const ks_u4 *lr = ring + (size_t)((l + NL - 1) & (NL - 1)) * 8;
				const ks_u4 cprev = lr[((r - 1) & 3) * 2], ccur = lr[(r & 3) * 2], bin = lr[(r & 3) * 2 + 1];
				ks_u4 co, bo;
				const bool zs = ks_tile_step_mixed<KIND, CIG>(P, c, ezs->ez, KS_T(l), r, r >= KS_FA(l) && r <= KS_FB(l), cprev, ccur, bin, k > 0 ? save + (size_t)(k - 1) * SW : save, co, bo,
				                                              KS_DIR(CIG) ? pbase + (size_t)k * prows : (ks_u4*)0, axH0, axT, axR);
				ring[(l * 4 + (r & 3)) * 2] = co; ring[(l * 4 + (r & 3)) * 2 + 1] = bo;
				if (r == KS_T(l).rb) save[(size_t)k * SW] = co;          // the block on the right reads it when this block has left the band (save_left)
				if (zs) ezs->done = 1;
			}
		KS_LANE_END }
		KS_SYNCWARP();
	}
	{ KS_LANE_LOOP(l) if (l == 0 && nd < c.ndiag && !ezs->done) { ezs->ez.zdropped = 1; ezs->ez.n_diag = nd; ezs->done = 1; } KS_LANE_END }
	KS_SYNCWARP();
#undef KS_T
#undef KS_K
#undef KS_ACT
#undef KS_FA
#undef KS_FB
}
