// ksw2_tile.cuh -- the B200 wavefront engine: one 16-lane block x a run of anti-diagonals ("tile").
//
// Design: the reference sweeps anti-diagonals r = i + j and, inside a diagonal, all
// target positions t of the band, 16 int8 lanes per SSE vector (ksw2_extz2_sse.c:101-289).  Here a
// *tile* is one 16-lane block k (lanes t = 16k..16k+15, the reference's vector granularity, which is
// observable) over consecutive diagonals.  All DP state of the block (u,v,x,y[,x2,y2],s,H) lives in
// REGISTERS for the whole tile; the only data crossing tile borders are
//   * one carry record per diagonal to the block on the right  (x,v[,x2] of lane 15 + H of lanes 13..15),
//   * one running arg-max record per diagonal (exact-max / Z-drop bookkeeping),
// both kept in small per-thread streams.  The dependency of a cell on (r-1, t-1) and (r-1, t) only
// (tex/ksw2.tex:146-152) makes any tile order legal that runs block k-1 before block k for a given
// diagonal, so a thread sweeps a panel of diagonals block by block without any inter-thread traffic.
//
// Bit-exactness notes: lanes outside the band but inside the 16-rounded range
// are evaluated exactly like the reference does (stale state included); the score row `s` is refreshed
// in unaligned 16-lane chunks starting at st0 (overshoot past en0, into the next block, is replayed
// when that block is first touched); arg-max ties resolve in the reference's 4-lane SIMD order.
#pragma once
#include "ksw2_prim.cuh"

#define KS_NEG_INF (-0x40000000)
enum { KS_Z = 0, KS_D = 1, KS_S = 2 };

#if defined(__CUDACC__)
typedef uint4 ks_u4;
#else
struct ks_u4 { uint32_t x, y, z, w; };
#endif
KS_HD ks_u4 ks_mk4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { ks_u4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// flag bits we need on the device (values: reference ksw2.h:8-18)
#define KSF_SCORE_ONLY 0x01
#define KSF_RIGHT      0x02
#define KSF_GENERIC_SC 0x04
#define KSF_APPROX_MAX 0x08
#define KSF_APPROX_DROP 0x10
#define KSF_EXTZ_ONLY  0x40
#define KSF_REV_CIGAR  0x80
#define KSF_SPLICE_FOR 0x100
#define KSF_SPLICE_REV 0x200
#define KSF_SPLICE_FLANK 0x400
#define KSF_EQX        0x800

struct KsParams {            // one per batch; prepared on the host (ksw2_host.cu: ks_prepare_params)
	int kind, flag, m;
	int q, e, q2, e2;        // gap costs; extd2: after the q+e <= q2+e2 swap (ksw2_extd2_sse.c:78)
	int h0sub;               // H[0] = v[0] - h0sub at r == 0 (extz2 2(q+e); extd2 pre-swap q+e; exts2 q+e)
	int qe_sub;              // per-diagonal term: extz2 keeps unsigned u,v and subtracts q+e; others 0
	int w;                   // caller's band (<0: none); exts2: -1
	int zdrop, zdrop_e, end_bonus;
	int long_thres, long_diff, e_far;
	int clamp;               // extz2: int8(mat[0]+2(q+e)); extd2: mat[0]; exts2: unused
	int init_a, init_b;      // initial u,v,x,y and x2,y2 (int8)
	int sz_init;             // initial s contribution: extz2 int8(2(q+e)), others 0
	int noncan, junc_bonus, semi;
	int smode;               // 0: 3-class LUT (match / mismatch / wildcard), 1: matrix lookup
	int gen_sc;              // KSW_EZ_GENERIC_SC semantics for the score-row write range
	int wild;                // wildcard code m-1
	int treload;             // saved blocks do not hold the coded target word (see ks_save_words)
	uint32_t lut_lo, lut_hi; // PRMT look-up: byte e = s-contribution of class e (0 eq, 1..3 ne, 4..7 wildcard)
	uint32_t tlow;           // low-nibble selector planted in the target bytes: 8|index of a LUT byte with msb 0
	int8_t zmat[32];         // smode 1 with m <= 5: unused; kept for alignment
	const int8_t *mat;       // smode 1: device pointer to an m*m table of s-contributions (already +2(q+e) for extz2)
};

struct KsPair {              // one alignment job as seen by a thread
	const uint8_t *query, *target, *junc;   // raw sequences (junc/target raw bytes are only read by exts2's splice set-up)
	const uint8_t *tenc;     // coded target, one 16-byte word per block, lanes in register order (ks_perm_pos); 16-byte aligned
	const uint8_t *qenc;     // coded REVERSED query qr[i] = code(query[qlen-1-i]), readable for i in [-KS_QPADL, qlen+KS_QPADR)
	int qlen, tlen, w, ndiag, tlen_;
};
#define KS_QPADL 32
#define KS_QPADR 64
KS_HD size_t ks_qenc_bytes(int qlen) { return (size_t)((qlen + KS_QPADL + KS_QPADR + 15) & ~15); }

struct KsEz {                // ksw_extz_t scalars (ksw2.h:33-42) held in registers during the fill
	int max, max_t, max_q, mqe, mqe_t, mte, mte_q, score, zdropped;
	int n_diag;              // diagonals the reference evaluates before it stops (cell accounting)
	int apx_H0, apx_t, apx_r;// KSW_EZ_APPROX_MAX: score and target position of the one tracked cell, next diagonal it has to be advanced on
};

// Stream records are 16-byte words (ks_u4) accessed with a stride so that the per-thread streams of a CTA can be
// interleaved in shared memory (conflict-free) on the device and contiguous (stride 1) in the host simulator:
//   carry  (block k -> block k+1, one per diagonal): x = x15 | v15<<8 | x2_15<<16 (raw int8 bytes), y,z,w = H[13],H[14],H[15]
//   best   (running SIMD-part arg-max of a diagonal): x = H, y = t (-1: none), z = H[st0] (for mqe)

KS_HD int ks_imax(int a, int b) { return a > b ? a : b; }
KS_HD int ks_imin(int a, int b) { return a < b ? a : b; }

// band geometry of diagonal r (ksw2_extz2_sse.c:105-116); returns false when the band is empty
KS_HD bool ks_geo(const KsPair &c, int r, int &st0, int &en0)
{
	st0 = ks_imax(ks_imax(0, r - c.qlen + 1), (r - c.w + 1) >> 1);
	en0 = ks_imin(ks_imin(c.tlen - 1, r), (r + c.w) >> 1);
	return st0 <= en0;
}
KS_HD int ks_rin(const KsPair &c, int k)  { return ks_imax(16 * k, 32 * k - c.w); }
KS_HD int ks_rout(const KsPair &c, int k) { return ks_imin(ks_imin(16 * k + 14 + c.qlen, 32 * k + 30 + c.w), c.ndiag - 1); }
// rows of direction bytes kept per block (uniform pitch)
KS_HD int ks_prows(int qlen, int tlen, int w) { return ks_imin(ks_imin(qlen + 15, 2 * w + 31), qlen + tlen - 1); }

KS_HD int ks_bnd(const KsParams &P, int r)   // first row/column value of u / v (extd2 :156-163, exts2 :188-195, extz2 :122-123)
{
	if (P.kind == KS_Z) return r ? P.q : 0;
	return r == 0 ? P.init_a : r < P.long_thres ? -P.e : r == P.long_thres ? P.long_diff : P.e_far;
}

// ---- Z-drop / max bookkeeping (ksw2.h:191-207) ----
KS_HD bool ks_zdrop(const KsParams &P, KsEz &ez, int H, int r, int t)
{
	if (H > ez.max) { ez.max = H; ez.max_t = t; ez.max_q = r - t; }
	else if (t >= ez.max_t && r - t >= ez.max_q) {
		int tl = t - ez.max_t, ql = (r - t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
		if (P.zdrop >= 0 && ez.max - H > P.zdrop + l * P.zdrop_e) { ez.zdropped = 1; return true; }
	}
	return false;
}

// ---- block state ----
template<int KIND> struct KsBlk {
	pk U[8], V[8], X[8], Y[8], SZ[8];
	pk X2[8];              // extd2 / exts2
	pk Y2[8];              // extd2: y2;  exts2: donor
	pk AC[8];              // exts2: acceptor
	int32_t H[16];
	uint32_t T[4], Q[4];   // class/code bytes, byte order per register j: lanes 2j, 2j+8, 2j+1, 2j+9
};
// 16-byte words of a saved block: carry, {T,Q} (2), int8 state arrays packed to bytes (1 word each), H (4)
template<int KIND> struct KsSaveWords { enum { value = 1 + 2 + (KIND == KS_Z ? 5 : KIND == KS_D ? 7 : 8) + 4 }; };
// KsParams::treload: the block's coded target word (which never changes) is not saved but re-read from the read-only coded target, one
// word less per saved block.  Pays when the saved state of all resident threads is about the size of L2 (short pairs: +4.7 % on the 150 bp
// workload), costs when it is far larger anyway (-3 % on 5 kb pairs): the host decides per batch (ks_save_words).
KS_HD int ks_save_words(const KsParams &P, int sw_full) { return sw_full - (P.treload ? 1 : 0); }
// pack / unpack one state array: 16 lanes (int8 << 8 in 8 registers) <-> 16 bytes in ks_perm_pos order
KS_HD ks_u4 ks_pack16(const pk *A) { return ks_mk4(prmt(A[0], A[1], 0x7531), prmt(A[2], A[3], 0x7531), prmt(A[4], A[5], 0x7531), prmt(A[6], A[7], 0x7531)); }
KS_HD void ks_unpack16(const ks_u4 w, pk *A)
{
	A[0] = prmt(w.x, 0u, 0x1404); A[1] = prmt(w.x, 0u, 0x3424); A[2] = prmt(w.y, 0u, 0x1404); A[3] = prmt(w.y, 0u, 0x3424);
	A[4] = prmt(w.z, 0u, 0x1404); A[5] = prmt(w.z, 0u, 0x3424); A[6] = prmt(w.w, 0u, 0x1404); A[7] = prmt(w.w, 0u, 0x3424);
}

// position of block lane L inside a 16-byte block word (T/Q code words and direction rows share it):
// bytes = lanes 0,8,1,9 | 2,10,3,11 | 4,12,5,13 | 6,14,7,15   (register j = lanes 2j, 2j+8, 2j+1, 2j+9)
KS_HD int ks_perm_pos(int L) { const int i = L & 7; return (i >> 1) * 4 + (i & 1) * 2 + (L >> 3); }

// sequence byte -> code byte stored in T[]/Q[] (class<<4 in LUT mode, raw code in matrix mode)
KS_HD uint32_t ks_code(const KsParams &P, int raw, bool is_target)
{
	if (P.smode) return (uint32_t)raw & 0xffu;
	const uint32_t cls = raw == P.wild ? 4u : ((uint32_t)raw & 3u);
	return (cls << 4) | (is_target ? P.tlow : 0u);
}
// one-off encoding of a pair (device: ks_encode_kernel; host simulator: same function).  i: target lane index in
// [0, tlen_*16) resp. reversed-query index in [-KS_QPADL, qlen + KS_QPADR); outside the sequence = code of 0 (the
// reference's zeroed padding, ksw2_extz2_sse.c:84,98-99)
KS_HD uint8_t ks_enc_t(const KsParams &P, const uint8_t *target, int tlen, int i) { return (uint8_t)ks_code(P, i < tlen ? target[i] : 0, true); }
KS_HD uint8_t ks_enc_q(const KsParams &P, const uint8_t *query, int qlen, int i) { return (uint8_t)ks_code(P, (i >= 0 && i < qlen) ? query[qlen - 1 - i] : 0, false); }

// dynamic lane reads without dynamic register indexing (keeps arrays in registers on the device); rare paths only
KS_HD int32_t ks_hget(const int32_t *H, int j)
{
#if defined(__CUDA_ARCH__)
	int32_t r = H[0];
#pragma unroll
	for (int i = 1; i < 16; ++i) r = (j == i) ? H[i] : r;
	return r;
#else
	return H[j];
#endif
}
template<int KIND> KS_HD int ks_uv(pk reg, int half) { return KIND == KS_Z ? lane_u(reg, half) : lane_s(reg, half); }

// lane masks of register i (lane i in the low half, lane i+8 in the high half): lanes >= L, lane == L
#if defined(__CUDACC__)
#define KS_MG(i, L) (((i) >= (L) ? 0x0000ffffu : 0u) | ((i) + 8 >= (L) ? 0xffff0000u : 0u))
#define KS_MGROW(L) { KS_MG(0, L), KS_MG(1, L), KS_MG(2, L), KS_MG(3, L), KS_MG(4, L), KS_MG(5, L), KS_MG(6, L), KS_MG(7, L) }
__constant__ uint32_t ks_mge_tab[17][8] = { KS_MGROW(0), KS_MGROW(1), KS_MGROW(2), KS_MGROW(3), KS_MGROW(4), KS_MGROW(5), KS_MGROW(6), KS_MGROW(7), KS_MGROW(8),
	KS_MGROW(9), KS_MGROW(10), KS_MGROW(11), KS_MGROW(12), KS_MGROW(13), KS_MGROW(14), KS_MGROW(15), KS_MGROW(16) };
#endif
KS_HD pk ks_maskge(int L, int i)
{
#if defined(__CUDA_ARCH__)
	return ks_mge_tab[L][i];
#else
	return ((i >= L) ? 0x0000ffffu : 0u) | ((i + 8 >= L) ? 0xffff0000u : 0u);
#endif
}
KS_HD int ks_clamp16(int v) { return v < 0 ? 0 : v > 16 ? 16 : v; }

// matrix look-up of the 16 lane scores (KSW_EZ_GENERIC_SC, m > 5, ...): the uncommon scoring mode, kept OUT OF LINE on the device so
// that its 16 loads and index arithmetic do not sit inside the step loops (instruction-cache footprint of the common path)
struct ks_pk8 { pk v[8]; };
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
static inline
#endif
ks_pk8 ks_score_mat(const int8_t *mat, int m, ks_u4 T, ks_u4 Q)
{
	ks_pk8 o;
	const uint32_t t[4] = {T.x, T.y, T.z, T.w}, q[4] = {Q.x, Q.y, Q.z, Q.w};
#pragma unroll
	for (int L = 0; L < 16; ++L) {
		const int pos = ks_perm_pos(L);
		const int a = (int)((t[pos >> 2] >> (8 * (pos & 3))) & 0xffu), b = (int)((q[pos >> 2] >> (8 * (pos & 3))) & 0xffu);
		const int s = mat[a * m + b];
		if (L < 8) o.v[L] = ((uint32_t)s & 0xffu) << 8; else o.v[L - 8] |= ((uint32_t)s & 0xffu) << 24;
	}
	return o;
}

// score-row refresh for diagonal r restricted to block lanes [lo, hi) (already clamped to 0..16)
template<int KIND> KS_HD void ks_score_row(const KsParams &P, KsBlk<KIND> &B, int lo, int hi)
{
	if (lo >= hi) return;
	pk nw[8];
	if (P.smode == 0) {
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t x = ks_class_bits(B.T[j], B.Q[j]);
			nw[2 * j]     = prmt(P.lut_lo, P.lut_hi, x);
			nw[2 * j + 1] = prmt(P.lut_lo, P.lut_hi, x >> 16);
		}
	} else {
		const ks_pk8 o = ks_score_mat(P.mat, P.m, ks_mk4(B.T[0], B.T[1], B.T[2], B.T[3]), ks_mk4(B.Q[0], B.Q[1], B.Q[2], B.Q[3]));
#pragma unroll
		for (int i = 0; i < 8; ++i) nw[i] = o.v[i];
	}
	if (lo == 0 && hi == 16) {
#pragma unroll
		for (int i = 0; i < 8; ++i) B.SZ[i] = nw[i];
	} else {
#pragma unroll
		for (int i = 0; i < 8; ++i) { const pk m = ks_maskge(lo, i) & ~ks_maskge(hi, i); B.SZ[i] = sel2(m, nw[i], B.SZ[i]); }
	}
}

// local write range of the score row on diagonal (st0,en0) for block base t0
KS_HD void ks_srange(const KsParams &P, int st0, int en0, int t0, int &lo, int &hi)
{
	const int wend = P.gen_sc ? en0 + 1 : st0 + 16 * ((en0 - st0) / 16 + 1);
	lo = ks_clamp16(st0 - t0); hi = ks_clamp16(wend - t0);
}

// slide the query window by one diagonal; nb = code byte of the query base entering lane 0
KS_HD void ks_qshift(uint32_t *Q, uint32_t nb)
{
	const uint32_t o3 = Q[3], o0 = Q[0];
	Q[3] = fshr16(Q[2], Q[3]); Q[2] = fshr16(Q[1], Q[2]); Q[1] = fshr16(o0, Q[1]);
	Q[0] = prmt(prmt(nb, o3, 0x0060u), o0, 0x5410u);        // bytes: nb.b0 | o3.b2 | o0.b0 | o0.b1  (two PRMT instead of five shift / mask ops)
}
// query window of diagonal r: lane L sees qr[qlen-1-r+t0+L]
template<int KIND> KS_HD void ks_qload(const KsPair &c, KsBlk<KIND> &B, int r, int t0)
{
	const uint8_t *p = c.qenc + (c.qlen - 1 - r + t0);
	uint32_t q[4] = {0u, 0u, 0u, 0u};
#pragma unroll
	for (int L = 0; L < 16; ++L) { const int pos = ks_perm_pos(L); q[pos >> 2] |= (uint32_t)p[L] << (8 * (pos & 3)); }
	B.Q[0] = q[0]; B.Q[1] = q[1]; B.Q[2] = q[2]; B.Q[3] = q[3];
}

// exts2 donor/acceptor values of target position t (ksw2_exts2_sse.c:119-171)
KS_HD void ks_splice(const KsParams &P, const KsPair &c, int t, int &don, int &acc)
{
	don = acc = 0;
	if (!(P.flag & (KSF_SPLICE_FOR | KSF_SPLICE_REV))) return;
	const bool fw = P.flag & KSF_SPLICE_FOR, rv = P.flag & KSF_SPLICE_REV, rc = P.flag & KSF_REV_CIGAR;
	const uint8_t *T = c.target, *jn = c.junc;
	const int tl = c.tlen;
	int8_t d = (int8_t)-P.noncan, a = (int8_t)-P.noncan;
	if (t < tl - 4) {
		int can = 0;
		if (fw && T[t + 1] == 2 && T[t + 2] == (rc ? 0 : 3)) can = 1;
		if (rv && T[t + 1] == 1 && T[t + 2] == (rc ? 0 : 3)) can = 1;
		if (can && (T[t + 3] == (rc ? 1 : 0) || T[t + 3] == (rc ? 3 : 2))) can = 2;
		if (can) d = (int8_t)(can == 2 ? 0 : P.semi);
	}
	if (jn && t < tl - 1) {
		const int bf = rc ? 2 : 1, br = rc ? 4 : 8;
		if ((fw && (jn[t + 1] & bf)) || (rv && (jn[t + 1] & br))) d = (int8_t)(d + P.junc_bonus);
	}
	if (t >= 2 && t < tl) {
		int can = 0;
		if (fw && T[t - 1] == (rc ? 3 : 0) && T[t] == 2) can = 1;
		if (rv && T[t - 1] == (rc ? 3 : 0) && T[t] == 1) can = 1;
		if (can && (T[t - 2] == (rc ? 0 : 1) || T[t - 2] == (rc ? 2 : 3))) can = 2;
		if (can) a = (int8_t)(can == 2 ? 0 : P.semi);
	}
	if (jn && t < tl) {
		const int bf = rc ? 1 : 2, br = rc ? 8 : 4;
		if ((fw && (jn[t] & bf)) || (rv && (jn[t] & br))) a = (int8_t)(a + P.junc_bonus);
	}
	don = d; acc = a;
}

// ---- exact-max bookkeeping -------------------------------------------------------------------------
// Per-diagonal context handed to the top block's finalisation
struct KsDiag { int r, st0, en0, en, en1, t0; bool is_first, have, qend; };

// arg-max over the SIMD-part lanes of a block in the reference's order (ksw2_extz2_sse.c:228-256):
// 4 SIMD lanes by (t - st0) % 4, inside a lane the first (lowest t) strict maximum, lanes merged in ascending order.
// Hm[]: H of the candidate lanes, INT_MIN+1 for lanes that do not take part.  Returns bT < 0 if there is no candidate.
#define KS_NOCAND (-0x7fffffff)
KS_HD int ks_block_max(const int32_t *Hm, int *m)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
	for (int a = 0; a < 4; ++a) m[a] = max(__vimax3_s32(Hm[a], Hm[a + 4], Hm[a + 8]), Hm[a + 12]);
	return max(__vimax3_s32(m[0], m[1], m[2]), m[3]);
#else
	for (int a = 0; a < 4; ++a) m[a] = ks_imax(ks_imax(Hm[a], Hm[a + 4]), ks_imax(Hm[a + 8], Hm[a + 12]));
	return ks_imax(ks_imax(m[0], m[1]), ks_imax(m[2], m[3]));
#endif
}
// which SIMD lane (bC) and target position (bT) hold the block maximum M (m[]: per-residue maxima from ks_block_max / ks_block_max_masked).
// H: the block's H, unmasked; mc: bit j set = lane j takes part.  Only the four lanes of the winning residue are inspected.
KS_HD void ks_block_arg(const int32_t *H, const int *m, int M, int st0, int t0, uint32_t mc, int &bT, int &bC)
{
	bT = -1; bC = 4;
	if (M == KS_NOCAND) return;
	const uint32_t E = (m[0] == M ? 1u : 0u) | (m[1] == M ? 2u : 0u) | (m[2] == M ? 4u : 0u) | (m[3] == M ? 8u : 0u);
	const int n0 = st0 & 3;                                  // residue (t & 3) of SIMD lane 0; t0 is a multiple of 16
	const uint32_t Er = ((E | (E << 4)) >> n0) & 15u;
	const int cl = (Er & 1u) ? 0 : (Er & 2u) ? 1 : (Er & 4u) ? 2 : 3;
	const int n = (n0 + cl) & 3;
	const int h0 = n == 0 ? H[0] : n == 1 ? H[1] : n == 2 ? H[2] : H[3];
	const int h1 = n == 0 ? H[4] : n == 1 ? H[5] : n == 2 ? H[6] : H[7];
	const int h2 = n == 0 ? H[8] : n == 1 ? H[9] : n == 2 ? H[10] : H[11];
	const uint32_t v = mc >> n;                              // bits 0, 4, 8, 12: validity of the residue's four lanes
	const int kq = (h0 == M && (v & 1u)) ? 0 : (h1 == M && (v & 0x10u)) ? 1 : (h2 == M && (v & 0x100u)) ? 2 : 3;
	bT = t0 + n + 4 * kq; bC = cl;
}
// per-residue and block maximum over the lanes whose bit is set in mc (KS_NOCAND if none)
KS_HD int ks_block_max_masked(const int32_t *H, uint32_t mc, int *m)
{
	int32_t Hm[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int j = 0; j < 16; ++j) Hm[j] = (mc & (1u << j)) ? H[j] : KS_NOCAND;
	return ks_block_max(Hm, m);
}
KS_HD void ks_block_argmax(const int32_t *Hm, int st0, int t0, int &bH, int &bT, int &bC)
{
	int m[4];
	bH = ks_block_max(Hm, m);
	ks_block_arg(Hm, m, bH, st0, t0, 0xffffu, bT, bC);
}

// Static-lane accessors for the block that holds en0 (lane J of the block): dispatched by a switch so that the register
// arrays are never indexed dynamically; everything else of the finalisation is shared, straight-line code.
template<int KIND, int J> KS_HD void ks_top_pre(const KsBlk<KIND> &B, int &h_own_prev, int &uvn_u, int &uvn_v0)
{
	h_own_prev = B.H[J > 0 ? J - 1 : 0];                                  // OLD H[en0-1] when it lives in this block
	uvn_u = ks_uv<KIND>(B.U[KS_REG(J)], KS_HALF(J));                      // new u[en0]
	uvn_v0 = ks_uv<KIND>(B.V[0], 0);                                      // new v[0] (only used when en0 == 0)
}
template<int KIND, int J> KS_HD void ks_top_post(const KsBlk<KIND> &B, int &h1, int &h2, int &h3)
{
	h1 = B.H[J >= 1 ? J - 1 : 0]; h2 = B.H[J >= 2 ? J - 2 : 0]; h3 = B.H[J >= 3 ? J - 3 : 0];   // updated H of the lanes below en0
}
template<int KIND, int J> KS_HD void ks_top_post_w(KsBlk<KIND> &B, int Hen0, int &h1, int &h2, int &h3)
{
	B.H[J] = Hen0;
	ks_top_post<KIND, J>(B, h1, h2, h3);
}
// H[J] = v for a run-time lane J, IN PLACE: sixteen predicated moves.  (A switch over J makes the compiler build a new copy of the
// whole 16-register array in every case: ~20 moves per step and ~340 instructions of code inside the step loop.)
KS_HD void ks_hset(int32_t *H, int J, int32_t v)
{
#if defined(__CUDA_ARCH__)
#define KS_HSET1(j) asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, " #j ";\n\t@p mov.b32 %0, %2;\n\t}" : "+r"(H[j]) : "r"(J), "r"(v));
	KS_HSET1(0) KS_HSET1(1) KS_HSET1(2) KS_HSET1(3) KS_HSET1(4) KS_HSET1(5) KS_HSET1(6) KS_HSET1(7)
	KS_HSET1(8) KS_HSET1(9) KS_HSET1(10) KS_HSET1(11) KS_HSET1(12) KS_HSET1(13) KS_HSET1(14) KS_HSET1(15)
#undef KS_HSET1
#else
	H[J] = v;
#endif
}
#define KS_SWITCH16(V, CALL) switch (V) { \
	case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
	case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; \
	case 8: CALL(8); break; case 9: CALL(9); break; case 10: CALL(10); break; case 11: CALL(11); break; \
	case 12: CALL(12); break; case 13: CALL(13); break; case 14: CALL(14); break; default: CALL(15); break; }

// ---- approximate-max mode -----------------------------------------------------------------------------
// KSW_EZ_APPROX_MAX (ksw2_extz2_sse.c:270-286, ksw2_extd2_sse.c:367-383, ksw2_exts2_sse.c:385-401; the reference's fast mode): no H[] row, no
// per-diagonal maximum, no mqe / mte; ONE cell is followed from (0,0): on each diagonal it stays at target position L (score += v[L]) or
// moves to L+1 (score += u[L+1]), whichever is larger, reading the NEW u / v of the diagonal.  L never decreases and (by induction over the
// diagonals) always ends inside [st0, en0], so the tracker fits the tile sweep: the block that owns the lane(s) to be read advances it on
// its step of that diagonal -- block L>>4, or its right neighbour when L is a lane 15 and L+1 is in the band (v[L] then comes from the
// left block's carry record of the same diagonal).  Blocks run left to right inside a panel and each runs its diagonals in ascending
// order, so the tracker, which only waits for blocks further right, is advanced through every diagonal of the panel in order.
// (H0, L, ar): the tracker before this step (ar = the diagonal it waits for); returns true when KSW_EZ_APPROX_DROP fired.
KS_HD pk ks_reg_dyn(const pk *A, int i)
{
#if defined(__CUDA_ARCH__)
	pk r = A[0];
#pragma unroll
	for (int n = 1; n < 8; ++n) r = (i == n) ? A[n] : r;
	return r;
#else
	return A[i];
#endif
}
template<int KIND> KS_HD int ks_uv_dyn(const pk *A, int L) { return ks_uv<KIND>(ks_reg_dyn(A, L & 7), (L >> 3) & 1); }

template<int KIND>
KS_HD bool ks_apx_step(const KsParams &P, const KsPair &c, KsEz &ez, const KsBlk<KIND> &B, int k, int r, int st0, int en0, uint32_t left_xv, int H0, int L, int ar)
{
	if (r != ar) return false;
	if (r == 0) { if (k != 0) return false; H0 = ks_uv<KIND>(B.V[0], 0) - P.h0sub; L = 0; }
	else if (L >= st0 && L < en0) {                           // L and L+1 in the band
		int d0, d1;
		if ((L >> 4) == k && (L & 15) != 15) { d0 = ks_uv_dyn<KIND>(B.V, L & 15); d1 = ks_uv_dyn<KIND>(B.U, (L & 15) + 1); }
		else if (((L + 1) >> 4) == k && (L & 15) == 15) { const int b = (int)((left_xv >> 8) & 0xffu); d0 = KIND == KS_Z ? b : (int)(int8_t)b; d1 = ks_uv<KIND>(B.U[0], 0); }
		else return false;                                    // another block's business
		d0 -= P.qe_sub; d1 -= P.qe_sub;
		if (d0 > d1) H0 += d0; else { H0 += d1; ++L; }
	} else if (L >= st0) {                                    // L == en0
		if ((L >> 4) != k) return false;
		H0 += ks_uv_dyn<KIND>(B.V, L & 15) - P.qe_sub;
	} else {                                                  // the band moved past L: L == st0 - 1
		if (((L + 1) >> 4) != k) return false;
		++L; H0 += ks_uv_dyn<KIND>(B.U, L & 15) - P.qe_sub;
	}
	ez.apx_H0 = H0; ez.apx_t = L; ez.apx_r = r + 1;
	if ((P.flag & KSF_APPROX_DROP) && (KIND != KS_Z || r > 0) && ks_zdrop(P, ez, H0, r, L)) { ez.n_diag = r + 1; return true; }
	if (r == c.ndiag - 1 && en0 == c.tlen - 1) ez.score = H0;
	return false;
}

// ---- one tile: block k over a run of diagonals, as begin / step / end --------------------------
// CIG: 0 score only, 1 left-aligned gaps, 2 right-aligned gaps (KSW_EZ_RIGHT); + 4: approximate-max mode (KSW_EZ_APPROX_MAX), a kernel variant of its
// own so that neither mode carries the other's code (the fill kernels are instruction-cache sensitive) or registers (no H[] in approximate mode)
#define KS_DIR(CIG) ((CIG) & 3)
#define KS_APX(CIG) (((CIG) >> 2) & 1)
// Two drivers use these pieces (ksw2_pair.cuh): one THREAD per alignment sweeping the blocks of a panel one after the other,
// and one WARP per alignment with the blocks of a panel spread over the lanes as a diagonal-skewed wavefront.
template<int KIND> struct KsTile {
	KsBlk<KIND> B;
	int k, t0, rin, ra, rb;
	pk INIT_A, INIT_B, CLAMP, QC1, Q2C1, NQE, NQE2;
	const uint8_t *qin;          // lane-0 code of diagonal r is qin[-r]
	const uint8_t *qp;           // running prefetch pointer: the steps run over consecutive diagonals, *qp is the code after qnext
	uint32_t qnext;              // prefetched code for the next diagonal
	ks_u4 last_out;
};

// Loads (or initialises) the state of block k for diagonals ra..rb.  seed = the carry record this block produced on the
// last diagonal before the panel (what the block on the right needs for its first diagonal).
template<int KIND>
KS_HD void ks_tile_begin(const KsParams &P, const KsPair &c, KsTile<KIND> &T, int k, int ra, int rb, const ks_u4 *save, ks_u4 &seed)
{
	KsBlk<KIND> &B = T.B;
	const int t0 = 16 * k, rin = ks_rin(c, k);
	T.k = k; T.t0 = t0; T.rin = rin; T.ra = ra; T.rb = rb;
	T.INIT_A = rep2(P.init_a); T.INIT_B = rep2(P.init_b);
	T.CLAMP = rep2(P.clamp); T.QC1 = rep2(P.q) + KS_ONE1; T.Q2C1 = rep2(P.q2) + KS_ONE1;
	T.NQE = rep2(-(P.q + P.e)); T.NQE2 = rep2(KIND == KS_D ? -(P.q2 + P.e2) : -P.q2);
	T.qin = c.qenc + (c.qlen - 1 + t0);
	T.last_out = ks_mk4(0u, (uint32_t)KS_NEG_INF, (uint32_t)KS_NEG_INF, (uint32_t)KS_NEG_INF);
	if (ra == rin) {                                   // first time this block is touched
		{ const ks_u4 tw = ((const ks_u4*)c.tenc)[k]; B.T[0] = tw.x; B.T[1] = tw.y; B.T[2] = tw.z; B.T[3] = tw.w; }
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			B.U[i] = B.V[i] = B.X[i] = B.Y[i] = T.INIT_A; B.SZ[i] = rep2(P.sz_init);
			if (KIND != KS_Z) B.X2[i] = T.INIT_B;
			if (KIND == KS_D) B.Y2[i] = T.INIT_B;
		}
#pragma unroll
		for (int j = 0; j < 16; ++j) B.H[j] = KS_NEG_INF;
		if (KIND == KS_S) {
#pragma unroll
			for (int i = 0; i < 8; ++i) B.Y2[i] = B.AC[i] = 0;
			for (int L = 0; L < 16; ++L) {
				int don, acc; ks_splice(P, c, t0 + L, don, acc);
				if (t0 + L >= c.tlen) { don = acc = (P.flag & (KSF_SPLICE_FOR | KSF_SPLICE_REV)) ? (int8_t)-P.noncan : 0; }   // memset tail
				const int i = KS_REG(L), h = KS_HALF(L);
#if defined(__CUDA_ARCH__)
#pragma unroll
				for (int n = 0; n < 8; ++n) if (n == i) { B.Y2[n] = set_lane(B.Y2[n], h, don); B.AC[n] = set_lane(B.AC[n], h, acc); }
#else
				B.Y2[i] = set_lane(B.Y2[i], h, don); B.AC[i] = set_lane(B.AC[i], h, acc);
#endif
			}
		}
		// replay score-row writes that overshot into this block before it became active
		const int r0 = ks_imax(0, rin - 31);
		ks_qload<KIND>(c, B, r0, t0);
		for (int r = r0; r < rin; ++r) {
			int st0, en0, lo, hi;
			if (!ks_geo(c, r, st0, en0)) break;
			ks_srange(P, st0, en0, t0, lo, hi);
			ks_score_row<KIND>(P, B, lo, hi);
			ks_qshift(B.Q, c.qenc[c.qlen - 1 - (r + 1) + t0]);
		}
		if (r0 < rin) ks_qload<KIND>(c, B, ra, t0);
		seed = T.last_out;
	} else {
		int wd = 0;
		seed = save[wd++];
		{ const ks_u4 a = P.treload ? ((const ks_u4*)c.tenc)[k] : save[wd++]; B.T[0] = a.x; B.T[1] = a.y; B.T[2] = a.z; B.T[3] = a.w; }
		{ const ks_u4 a = save[wd++]; B.Q[0] = a.x; B.Q[1] = a.y; B.Q[2] = a.z; B.Q[3] = a.w; }
#define KS_LD(ARR) ks_unpack16(save[wd++], ARR);
		KS_LD(B.U) KS_LD(B.V) KS_LD(B.X) KS_LD(B.Y) KS_LD(B.SZ)
		if (KIND != KS_Z) { KS_LD(B.X2) KS_LD(B.Y2) }
		if (KIND == KS_S) { KS_LD(B.AC) }
#undef KS_LD
#pragma unroll
		for (int j = 0; j < 4; ++j) { ks_u4 a = save[wd++]; B.H[4 * j] = (int32_t)a.x; B.H[4 * j + 1] = (int32_t)a.y; B.H[4 * j + 2] = (int32_t)a.z; B.H[4 * j + 3] = (int32_t)a.w; }
		// the saved window belongs to the last diagonal of the previous panel (ra - 1): advance it
		ks_qshift(B.Q, T.qin[-ra]);
	}
	T.qnext = T.qin[-(ra + 1)];
	T.qp = T.qin - (ra + 1);
}

// start of a step: slide the query window to diagonal r and keep the prefetch of the coming codes going
template<int KIND> KS_HD void ks_qadvance(KsTile<KIND> &T, int r)
{
	if (r > T.ra) ks_qshift(T.B.Q, T.qnext);
	T.qnext = *T.qp--;                                  // qin[-(r + 1)]  (a lead of two diagonals was measured: +0.5 % / -0.6 %, not kept)
}

// carry word of a block for the block on its right: bytes 0..2 = x, v, x2 of lane 15 (the high bytes of register 7), byte 3 = 0.  The zero
// bytes come from the low byte of a lane, which is zero in every state register.
template<int KIND> KS_HD uint32_t ks_carry_word(const KsBlk<KIND> &B)
{
	const uint32_t t = prmt(B.X[7], B.V[7], 0x4473u);
	return KIND != KS_Z ? prmt(t, B.X2[7], 0x2710u) : t;
}

// The recurrence itself for all 16 lanes of the block on one diagonal (ksw2_extz2_sse.c:145-223, ksw2_extd2_sse.c:177-317,
// ksw2_exts2_sse.c:207-330).  xv: carry word, bytes 0..2 = x, v, x2 of target position t0-1 (what lane 0 reads); D: direction bytes (<< 8) if CIG.
template<int KIND, int CIG>
KS_HD void ks_core(KsTile<KIND> &T, uint32_t xv, bool quirk_x, bool quirk_v, pk *D)
{
	KsBlk<KIND> &B = T.B;
	// lane 0 reads (x, v, x2) of target position t0-1 = bytes 0, 1, 2 of the carry word xv; lane 8 reads lane 7.  One PRMT each: byte 1 <- the
	// carry byte, byte 3 <- lane 7's value, bytes 0 and 2 <- the (always zero) low byte of lane 7
	pk px  = prmt(xv, B.X[7], 0x5404u);
	pk pv  = prmt(xv, B.V[7], 0x5414u);
	pk px2 = KIND != KS_Z ? prmt(xv, B.X2[7], 0x5424u) : 0u;
	const pk qmx = quirk_x ? 0x0000ff00u : 0u, qmv = quirk_v ? 0x0000ff00u : 0u;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		pk xt = px, vt = pv, x2t = px2;
		if (KIND == KS_Z && i >= 1 && i <= 3) { xt |= qmx; vt |= qmv; }
		px = B.X[i]; pv = B.V[i]; if (KIND != KS_Z) px2 = B.X2[i];
		const pk ut = B.U[i];
		pk a = add2(xt, vt), b = add2(B.Y[i], ut), z = B.SZ[i], d = 0;
		if (KIND == KS_Z) {
			if (CIG == 0) z = maxs2(z, a);
			else if (CIG == 1) {
				const pk z1 = maxs2(z, a); d = nz_one2(z1 ^ z);
				const pk z2 = maxs2(z1, b), m = nz_one2(z2 ^ z1);
				d = maxu2(d, add2(m, m)); z = z1;
			} else {
				d = nz_one2(mins2(a, z) ^ z) ^ 0x01000100u;             // !(z > a)
				const pk z1 = maxs2(z, a), m = nz_one2(mins2(b, z1) ^ z1) ^ 0x01000100u;   // !(z1 > b)
				d = maxu2(d, add2(m, m)); z = z1;
			}
			z = maxu2(z, b); z = minu2(z, T.CLAMP);
		} else {
			pk a2 = add2(x2t, vt), v3, v4 = 0;
			if (KIND == KS_D) { v3 = a2; v4 = add2(B.Y2[i], ut); } else v3 = add2(a2, B.AC[i]);
			if (CIG == 0) {
				z = max3s2(z, a, b);
				if (KIND == KS_D) z = max3s2(z, v3, v4); else z = maxs2(z, v3);
			} else {
				// Value and direction in ONE 16-bit maximum: every candidate carries its index in the (otherwise zero) low byte of its
				// lane, arranged so that a tie goes to the EARLIER candidate when gaps are left-aligned (the reference tests '>' in the
				// order s, a, b, a2, b2: ksw2_extd2_sse.c:240-262) and to the LATER one with KSW_EZ_RIGHT (:275-297).
				const uint32_t NC = KIND == KS_D ? 4u : 3u, ONE = 0x00010001u;
				const pk t0 = CIG == 1 ? NC * ONE : 0u, t1 = CIG == 1 ? (NC - 1) * ONE : ONE, t2 = CIG == 1 ? (NC - 2) * ONE : 2u * ONE;
				const pk t3 = CIG == 1 ? (NC - 3) * ONE : 3u * ONE, t4 = CIG == 1 ? 0u : 4u * ONE;
				pk zt = max3s2(z | t0, a | t1, b | t2);
				zt = KIND == KS_D ? max3s2(zt, v3 | t3, v4 | t4) : maxs2(zt, v3 | t3);
				const pk tg = zt & 0x00070007u;
				z = zt & 0xff00ff00u;
				d = CIG == 1 ? NC * 0x01000100u - (tg << 8) : (tg << 8);
			}
			if (KIND == KS_D) z = mins2(z, T.CLAMP);
			// second-piece / intron state
			const pk nzq2 = add2(not2(z), T.Q2C1);                        // q2 - z
			const pk a2p = add2(a2, nzq2);                              // a2 - (z - q2)
			if (KIND == KS_D) {
				const pk b2p = add2(v4, nzq2);
				const pk mx = maxs2(a2p, 0), my = maxs2(b2p, 0);
				B.X2[i] = add2(mx, T.NQE2); B.Y2[i] = add2(my, T.NQE2);
				if (CIG == 1) d += nz_one2(mx) * 0x20u + nz_one2(my) * 0x40u;
				if (CIG == 2) d += ((~a2p & 0x80008000u) >> 2) + ((~b2p & 0x80008000u) >> 1);
			} else {
				const pk don = B.Y2[i], mx = maxs2(a2p, don);
				B.X2[i] = add2(mx, T.NQE2);
				if (CIG == 1) d += nz_one2(mx ^ don) * 0x20u;                               // a2 > donor
				if (CIG == 2) d += (nz_one2(mins2(a2p, don) ^ don) ^ 0x01000100u) * 0x20u;   // !(donor > a2)
			}
		}
		const pk zp = plus_one2(z), nzq = add2(not2(z), T.QC1);                   // q - z  (exact: ~z + q + 1/256)
		B.U[i] = add2(zp, not2(vt)); B.V[i] = add2(zp, not2(ut));
		pk mx, my;
		if (CIG == 2) {
			const pk ap = add2(a, nzq), bp = add2(b, nzq);                       // a - (z - q), b - (z - q)
			mx = maxs2(ap, 0); my = maxs2(bp, 0);
			d += ((~ap & 0x80008000u) >> 4) + ((~bp & 0x80008000u) >> 3);
		} else { mx = addmax0s2(a, nzq); my = addmax0s2(b, nzq); }          // max(a - (z - q), 0): one VIADDMNMX each
		if (KIND == KS_Z) { B.X[i] = mx; B.Y[i] = my; } else { B.X[i] = add2(mx, T.NQE); B.Y[i] = add2(my, T.NQE); }
		if (CIG == 1) d += nz_one2(mx) * 0x08u + nz_one2(my) * 0x10u;
		D[i] = d;
	}
}
KS_HD ks_u4 ks_pack_dirs(const pk *D) { return ks_mk4(prmt(D[0], D[1], 0x7531), prmt(D[2], D[3], 0x7531), prmt(D[4], D[5], 0x7531), prmt(D[6], D[7], 0x7531)); }

// One diagonal r of the tile, in three pieces so that the warp-cooperative driver can run lanes with different step kinds through ONE
// (converged) copy of the recurrence: ks_step_pre (band geometry, carry-in of lane 0, boundary lane, score row), ks_core, ks_step_post
// (exact max / approximate tracker, finalisation of the diagonal, records).
//   cprev / ccur: carry records of the block on the left for diagonals r-1 / r; bin: its arg-max record for diagonal r;
//   save_left: the left block's persisted record (used when it was not evaluated on r-1).
// The post step writes this block's records for diagonal r to cout / bout and returns true when Z-drop fired (ez.n_diag set).
struct KsStepCtx { int st0, en0, en; bool is_first, is_top, have, quirk_x, quirk_v; uint32_t xv; };

template<int KIND>
KS_HD void ks_step_pre(const KsParams &P, const KsPair &c, KsTile<KIND> &T, int r, const ks_u4 cprev, KsStepCtx &S)
{
	const int k = T.k, t0 = T.t0;
	int st0, en0;
	ks_geo(c, r, st0, en0);                           // non-empty by construction of the panel
	const int st = st0 & ~15, en = en0 | 15;
	const bool is_first = (st == t0), is_top = ((en0 >> 4) == k);
	ks_qadvance<KIND>(T, r);                            // slide the query window; prefetch the next code(s) (the coded query is padded on both sides)

	// was the block on the left evaluated on diagonal r-1?  (else its values are older: "last_st/last_en" test, :119)
	// (a block that does not hold st0 always has a live left neighbour on r-1: st(r-1) <= st(r) < t0 and en0(r-1) >= en0(r) - 1 >= t0 - 1)
	bool have = k > 0 && r > 0;
	if (have && is_first) {
		int pst0, pen0;
		have = ks_geo(c, r - 1, pst0, pen0) && (t0 - 1 >= (pst0 & ~15)) && (t0 - 1 <= (pen0 | 15));
	}
	// ---- carry-in for lane 0 ----
	int cx, cv, cx2;
	bool quirk_v = false, quirk_x = false;
	if (is_first) {
		if (k > 0) {
			if (have) { const uint32_t xv = cprev.x; cx = (int8_t)(xv & 0xff); cv = (int8_t)((xv >> 8) & 0xff); cx2 = (int8_t)((xv >> 16) & 0xff); }
			else { cx = P.init_a; cv = P.init_a; cx2 = P.init_b; }
		} else { cx = P.init_a; cx2 = P.init_b; cv = ks_bnd(P, r); }
		if (KIND == KS_Z) { cx = (int8_t)cx; cv = (int8_t)cv; quirk_x = cx < 0; quirk_v = cv < 0; }   // ksw2_extz2_sse.c:146-147 sign-extending move
	} else {
		const uint32_t xv = cprev.x; cx = (int8_t)(xv & 0xff); cv = (int8_t)((xv >> 8) & 0xff); cx2 = (int8_t)((xv >> 16) & 0xff);
	}
	// ---- first-row boundary lane t == r (:123) ----
		if (is_top && (r >> 4) == k) {
			const int j = r & 15;
			const pk bu = rep2(ks_bnd(P, r));
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const pk m = ks_maskge(j, i) & ~ks_maskge(j + 1, i);
				T.B.Y[i] = sel2(m, T.INIT_A, T.B.Y[i]); T.B.U[i] = sel2(m, bu, T.B.U[i]);
				if (KIND == KS_D) T.B.Y2[i] = sel2(m, T.INIT_B, T.B.Y2[i]);
			}
		}
		// ---- score row ----
		// (a block that starts at or above st0 and lies entirely below en0's block is always covered completely: the write range
		//  [st0, st0 + 16*((en0-st0)/16 + 1)) ends above en0)
		if (st0 <= t0 && !is_top) ks_score_row<KIND>(P, T.B, 0, 16);
		else { int lo, hi; ks_srange(P, st0, en0, t0, lo, hi); ks_score_row<KIND>(P, T.B, lo, hi); }

	S.st0 = st0; S.en0 = en0; S.en = en; S.is_first = is_first; S.is_top = is_top; S.have = have; S.quirk_x = quirk_x; S.quirk_v = quirk_v;
	S.xv = ((uint32_t)cx & 0xffu) | (((uint32_t)cv & 0xffu) << 8) | (((uint32_t)cx2 & 0xffu) << 16);
}

template<int KIND, int CIG>
KS_HD bool ks_step_post(const KsParams &P, const KsPair &c, KsEz &ez, KsTile<KIND> &T, int r, const KsStepCtx &S, const ks_u4 cprev, const ks_u4 ccur, const ks_u4 bin,
                        const ks_u4 *save_left, ks_u4 &cout, ks_u4 &bout, int axH0, int axT, int axR)
{
	const int k = T.k, t0 = T.t0, st0 = S.st0, en0 = S.en0, en = S.en;
	const bool is_first = S.is_first, is_top = S.is_top, have = S.have;
	if (KS_APX(CIG)) {                                            // approximate-max mode: no H[], no per-diagonal maximum (ks_apx_step)
		const bool stop_a = ks_apx_step<KIND>(P, c, ez, T.B, k, r, st0, en0, ccur.x, axH0, axT, axR);
		bout = ks_mk4((uint32_t)KS_NOCAND, (uint32_t)-1, (uint32_t)KS_NEG_INF, 0u);
		cout = ks_mk4(ks_carry_word<KIND>(T.B), 0u, 0u, 0u);
		T.last_out = cout;
		return stop_a;
	}
	// ---- exact max: H[], per-diagonal arg-max in the reference's SIMD order (:224-269) ----
	const int lo = st0 - t0;                                      // first in-band lane of this block (may be < 0)
	const int hi = is_top ? en0 - t0 : 16;                        // lanes [lo, hi) get H[t] += v[t] - qe; lane hi (top block) is en0
	const int en1 = st0 + (en0 - st0) / 4 * 4, e1 = en1 - t0;     // SIMD part is [st0, en1), scalar tail [en1, en0)
	const bool qend = (r - st0 == c.qlen - 1);
	bool stop = false;
	int Hen0 = 0, h1 = 0, h2 = 0, h3 = 0;
	if (r == 0) { T.B.H[0] = ks_uv<KIND>(T.B.V[0], 0) - P.h0sub; Hen0 = T.B.H[0]; }   // only block 0, en0 == 0
	else {
		if (is_top) {                                             // H[en0] first, from the OLD H[en0-1] (:226)
			int h_own = 0, uvn_u = 0, uvn_v0 = 0;
#define KS_CALL(J) ks_top_pre<KIND, J>(T.B, h_own, uvn_u, uvn_v0)
			KS_SWITCH16(hi, KS_CALL)
#undef KS_CALL
			int hprev, uvn;
			if (hi > 0) { hprev = h_own; uvn = uvn_u; }
			else if (en0 > 0) { hprev = have ? (int32_t)cprev.w : (int32_t)save_left[0].w; uvn = uvn_u; }   // H[16k-1]: live or last persisted
			else { hprev = h_own; uvn = uvn_v0; }
			Hen0 = hprev + uvn - P.qe_sub;
		}
		if (lo <= 0) {
			// every lane: lanes above en0 carry no information yet (a lane is assigned when it first becomes en0, ks_top_post) and
			// here no lane of the block lies below st0
#pragma unroll
			for (int j = 0; j < 16; ++j) T.B.H[j] += ks_uv<KIND>(T.B.V[KS_REG(j)], KS_HALF(j)) - P.qe_sub;
		} else {
			const uint32_t mu = (0xffffu << lo) & ~(0xffff0000u >> (16 - hi));        // lanes [lo, hi), 0 < lo <= 15, 0 <= hi <= 16
#pragma unroll
			for (int j = 0; j < 16; ++j) if (mu & (1u << j)) T.B.H[j] += ks_uv<KIND>(T.B.V[KS_REG(j)], KS_HALF(j)) - P.qe_sub;
		}
		if (is_top) {
			// H[en0] = Hen0.  extz2 kernels: in place (ks_hset) + a read-only switch; the dual-gap / splice kernels keep the write inside the
			// switch -- measured: with ks_hset ptxas trades their 186 registers for 168 + spills and the 5 kb CIGAR workload loses 4.6 %
			if (KIND == KS_Z) {
				ks_hset(T.B.H, hi, Hen0);
#define KS_CALL(J) ks_top_post<KIND, J>(T.B, h1, h2, h3)
				KS_SWITCH16(hi, KS_CALL)
#undef KS_CALL
			} else {
#define KS_CALL(J) ks_top_post_w<KIND, J>(T.B, Hen0, h1, h2, h3)
				KS_SWITCH16(hi, KS_CALL)
#undef KS_CALL
			}
		}
	}
	// block maximum over the SIMD-part lanes [lo, e1) (all below en0); its position is only worked out if it can beat the left blocks
	int bH = KS_NOCAND, bT = -1, bC = 4, hst0 = KS_NEG_INF;
	if (r > 0) {
		int m4[4];
		uint32_t mc = 0xffffu;                                    // candidate lanes [lo, e1)
		if (lo <= 0 && e1 >= 16) bH = ks_block_max(T.B.H, m4);
		else {
			const int lc = ks_clamp16(lo), ec = ks_clamp16(e1);
			mc = ec > lc ? (0xffffu << lc) & (0xffffu >> (16 - ec)) : 0u;
			bH = ks_block_max_masked(T.B.H, mc, m4);
		}
		if (!is_first) {
			const int sH = (int32_t)bin.x, sT = (int32_t)bin.y;
			if (sT < 0 || bH >= sH) ks_block_arg(T.B.H, m4, bH, st0, t0, mc, bT, bC);
			if (sT >= 0) {
				const int sC = (sT - st0) & 3;
				if (bT < 0 || sH > bH || (sH == bH && sC <= bC)) { bH = sH; bT = sT; }
			}
			hst0 = (int32_t)bin.z;
		} else ks_block_arg(T.B.H, m4, bH, st0, t0, mc, bT, bC);
	}
	if (is_first && qend) hst0 = ks_hget(T.B.H, lo);
	if (!is_top) bout = ks_mk4((uint32_t)bH, (uint32_t)bT, (uint32_t)hst0, 0u);
	else {
		// ---- finalise diagonal r (:226-269) ----
		int max_H = Hen0, max_t = en0;
		if (r > 0) {
			if (bT >= 0 && bH > max_H) { max_H = bH; max_t = bT; }
			// scalar tail [en1, en0): up to three lanes, possibly in the block on the left (its records of this diagonal: ccur)
#pragma unroll
			for (int d = 3; d >= 1; --d) {
				const int t = en0 - d;
				if (t >= en1) {
					int ht;
					if (hi - d >= 0) ht = d == 1 ? h1 : d == 2 ? h2 : h3;
					else { const int dd = t0 - t; ht = (int32_t)(dd == 1 ? ccur.w : dd == 2 ? ccur.z : ccur.y); }
					if (ht > max_H) { max_H = ht; max_t = t; }
				}
			}
		} else max_t = 0;
		if (en0 == c.tlen - 1 && Hen0 > ez.mte) { ez.mte = Hen0; ez.mte_q = r - en; }
		if (qend && hst0 > ez.mqe) { ez.mqe = hst0; ez.mqe_t = st0; }
		if (ks_zdrop(P, ez, max_H, r, max_t)) { ez.n_diag = r + 1; stop = true; }
		else if (r == c.ndiag - 1 && en0 == c.tlen - 1) ez.score = Hen0;
		bout = ks_mk4((uint32_t)KS_NOCAND, (uint32_t)-1, (uint32_t)KS_NEG_INF, 0u);
	}
	cout = ks_mk4(ks_carry_word<KIND>(T.B),
	              (uint32_t)T.B.H[13], (uint32_t)T.B.H[14], (uint32_t)T.B.H[15]);
	T.last_out = cout;
	return stop;
}

template<int KIND, int CIG>
KS_HD bool ks_tile_step(const KsParams &P, const KsPair &c, KsEz &ez, KsTile<KIND> &T, int r, const ks_u4 cprev, const ks_u4 ccur, const ks_u4 bin,
                        const ks_u4 *save_left, ks_u4 &cout, ks_u4 &bout, ks_u4 *prow, int axH0, int axT, int axR)
{
	KsStepCtx S;
	ks_step_pre<KIND>(P, c, T, r, cprev, S);
	pk D[8];
	ks_core<KIND, KS_DIR(CIG)>(T, S.xv, S.quirk_x, S.quirk_v, D);
	if (KS_DIR(CIG)) prow[r - T.rin] = ks_pack_dirs(D);
	return ks_step_post<KIND, CIG>(P, c, ez, T, r, S, cprev, ccur, bin, save_left, cout, bout, axH0, axT, axR);
}

// One diagonal r of a block STRICTLY INSIDE the band: st0 < t0 (the block on the left is live on r-1 and r) and en0 >= t0 + 19 (all 16
// lanes lie below en0 and inside the SIMD part [st0, en1) of the arg-max, :228-256).  Then there is no boundary lane, no partial
// score row, no stale-neighbour test, no finalisation: the step is the recurrence, H += v - qe, the block maximum and two records.
// Same results as ks_tile_step() on such a diagonal (which stays the reference point; the host simulator fuzzes both).
template<int KIND, int CIG>
KS_HD bool ks_step_post_fast(const KsParams &P, const KsPair &c, KsEz &ez, KsTile<KIND> &T, int r, int st0, uint32_t ccur_xv, const ks_u4 bin,
                             ks_u4 &cout, ks_u4 &bout, int axH0, int axT, int axR)
{
	KsBlk<KIND> &B = T.B;
	if (KS_APX(CIG)) {
		const bool stop_a = ks_apx_step<KIND>(P, c, ez, B, T.k, r, st0, ks_imin(ks_imin(c.tlen - 1, r), (r + c.w) >> 1), ccur_xv, axH0, axT, axR);
		bout = bin;
		cout = ks_mk4(ks_carry_word<KIND>(B), 0u, 0u, 0u);
		T.last_out = cout;
		return stop_a;
	}
#pragma unroll
	for (int j = 0; j < 16; ++j) B.H[j] += ks_uv<KIND>(B.V[KS_REG(j)], KS_HALF(j)) - P.qe_sub;
	const int sH = (int32_t)bin.x, sT = (int32_t)bin.y;
#ifndef KS_NO_ARG_KEYS
	// Maximum AND its position from one max
	// tree over keys  H*16 + 4*(3 - SIMD lane) + (3 - quarter)  -- the reference's tie order (lower SIMD lane, then lower t, :228-256) is
	// the order of the low four bits, so no data-dependent branch and no select chains (ks_block_arg runs on 2/3 of the interior steps).
	// |H| < 2^27 on every lane of an interior block (they all are, or were, real cells).  Measured (round 2, B200): +4 % on the 150 bp
	// workload, +3.7 % on the 5 kb CIGAR workload over the branchy select chain below (profiles/r2_ab.txt).
	int bH, bT, bC;
	{
		const int n0 = st0 & 3;
		int m1[4];
#pragma unroll
		for (int a = 0; a < 4; ++a) {
			const int k0 = B.H[a] * 16 + 3, k1 = B.H[a + 4] * 16 + 2, k2 = B.H[a + 8] * 16 + 1, k3 = B.H[a + 12] * 16;
			m1[a] = ks_imax(ks_imax(k0, k1), ks_imax(k2, k3)) + 4 * (3 - ((a - n0) & 3));
		}
		const int km = ks_imax(ks_imax(m1[0], m1[1]), ks_imax(m1[2], m1[3]));
		bH = km >> 4; bC = 3 - ((km >> 2) & 3);
		bT = T.t0 + ((n0 + bC) & 3) + 4 * (3 - (km & 3));
	}
	if (sT >= 0) {
		const int sC = (sT - st0) & 3;
		if (sH > bH || (sH == bH && sC <= bC)) { bH = sH; bT = sT; }
	}
#else
	int m4[4], bT = -1, bC = 4;
	int bH = ks_block_max(B.H, m4);
	if (sT < 0 || bH >= sH) ks_block_arg(B.H, m4, bH, st0, T.t0, 0xffffu, bT, bC);
	if (sT >= 0) {
		const int sC = (sT - st0) & 3;
		if (bT < 0 || sH > bH || (sH == bH && sC <= bC)) { bH = sH; bT = sT; }
	}
#endif
	bout = ks_mk4((uint32_t)bH, (uint32_t)bT, bin.z, 0u);
	cout = ks_mk4(ks_carry_word<KIND>(B),
	              (uint32_t)B.H[13], (uint32_t)B.H[14], (uint32_t)B.H[15]);
	T.last_out = cout;
	return false;
}

template<int KIND, int CIG>
KS_HD bool ks_tile_step_fast(const KsParams &P, const KsPair &c, KsEz &ez, KsTile<KIND> &T, int r, int st0, const ks_u4 cprev, uint32_t ccur_xv, const ks_u4 bin,
                             ks_u4 &cout, ks_u4 &bout, ks_u4 *prow, int axH0, int axT, int axR)
{
	ks_qadvance<KIND>(T, r);
	ks_score_row<KIND>(P, T.B, 0, 16);
	pk D[8];
	ks_core<KIND, KS_DIR(CIG)>(T, cprev.x, false, false, D);
	if (KS_DIR(CIG)) prow[r - T.rin] = ks_pack_dirs(D);
	return ks_step_post_fast<KIND, CIG>(P, c, ez, T, r, st0, ccur_xv, bin, cout, bout, axH0, axT, axR);
}

// The same diagonal for a WARP whose lanes need different step kinds (warp-cooperative driver, banded pairs: the lanes at the band's two
// edges take the general step, the others the interior one).  Calling ks_tile_step / ks_tile_step_fast under a branch would run the
// recurrence twice, once per branch side; here only the small pre / post pieces diverge and ks_core is executed once, by all lanes together.
template<int KIND, int CIG>
KS_HD bool ks_tile_step_mixed(const KsParams &P, const KsPair &c, KsEz &ez, KsTile<KIND> &T, int r, bool interior, const ks_u4 cprev, const ks_u4 ccur, const ks_u4 bin,
                              const ks_u4 *save_left, ks_u4 &cout, ks_u4 &bout, ks_u4 *prow, int axH0, int axT, int axR)
{
	KsStepCtx S;
	if (interior) {
		ks_qadvance<KIND>(T, r);
		ks_score_row<KIND>(P, T.B, 0, 16);
		S.st0 = ks_imax(ks_imax(0, r - c.qlen + 1), (r - c.w + 1) >> 1); S.xv = cprev.x; S.quirk_x = S.quirk_v = false;
		S.en0 = S.en = 0; S.is_first = S.is_top = S.have = false;
	} else ks_step_pre<KIND>(P, c, T, r, cprev, S);
	pk D[8];
	ks_core<KIND, KS_DIR(CIG)>(T, S.xv, S.quirk_x, S.quirk_v, D);
	if (KS_DIR(CIG)) prow[r - T.rin] = ks_pack_dirs(D);
	if (interior) return ks_step_post_fast<KIND, CIG>(P, c, ez, T, r, S.st0, ccur.x, bin, cout, bout, axH0, axT, axR);
	return ks_step_post<KIND, CIG>(P, c, ez, T, r, S, cprev, ccur, bin, save_left, cout, bout, axH0, axT, axR);
}


// diagonals [fa, fb] of the tile (k, ra..rb) on which the block is strictly inside the band, i.e. ks_tile_step_fast applies:
// en0(r) >= t0 + 19 and st0(r) < t0; fa > fb if there are none
KS_HD void ks_fast_range(const KsPair &c, int k, int ra, int rb, int &fa, int &fb)
{
	fa = rb + 1; fb = rb;
#ifndef KS_NO_FAST_STEP
	if (k > 0 && c.tlen - 1 >= 16 * k + 19) {
		const int X = 16 * k + 19;
		fa = ks_imax(ra, ks_imax(X, 2 * X - c.w));
		fb = ks_imin(rb, ks_imin(16 * k + c.qlen - 2, 32 * k + c.w - 2));
		if (fa > fb) fa = rb + 1;
	}
#endif
}

// Persists the block: the last carry record always (the block on the right may still need it), the full state only if the
// block has diagonals left after rb.
template<int KIND>
KS_HD void ks_tile_end(const KsParams &P, const KsPair &c, KsTile<KIND> &T, ks_u4 *save)
{
	KsBlk<KIND> &B = T.B;
	int wd = 0;
	save[wd++] = T.last_out;
	if (T.rb < ks_rout(c, T.k)) {
		if (!P.treload) save[wd++] = ks_mk4(B.T[0], B.T[1], B.T[2], B.T[3]);
		save[wd++] = ks_mk4(B.Q[0], B.Q[1], B.Q[2], B.Q[3]);
#define KS_ST(ARR) save[wd++] = ks_pack16(ARR);
		KS_ST(B.U) KS_ST(B.V) KS_ST(B.X) KS_ST(B.Y) KS_ST(B.SZ)
		if (KIND != KS_Z) { KS_ST(B.X2) KS_ST(B.Y2) }
		if (KIND == KS_S) { KS_ST(B.AC) }
#undef KS_ST
#pragma unroll
		for (int j = 0; j < 4; ++j) save[wd++] = ks_mk4((uint32_t)B.H[4 * j], (uint32_t)B.H[4 * j + 1], (uint32_t)B.H[4 * j + 2], (uint32_t)B.H[4 * j + 3]);
	}
}

// Thread-per-alignment driver piece: the whole tile in one go.  Streams live in (shared) memory with element stride sst:
// cs: carry records indexed by (r - R + 1), updated IN PLACE (the left block's record of diagonal r is read, then replaced by
// this block's; the previous one travels in a register); best: arg-max records indexed by (r - R), also in place.
template<int KIND, int CIG>
KS_HD void ks_tile(const KsParams &P, const KsPair &c, KsEz &ez, int k, int ra, int rb, int R,
                   ks_u4 *save, const ks_u4 *save_left, ks_u4 *cs, ks_u4 *best, int sst, ks_u4 *prow, bool &done)
{
	KsTile<KIND> T;
	ks_u4 seed;
	ks_u4 cprev = cs[(size_t)(ra - R) * sst];                  // the left block's record of diagonal ra-1 (read before slot 0 is re-used)
	ks_tile_begin<KIND>(P, c, T, k, ra, rb, save, seed);
	// (asking L2 for the next block's slot a tile ahead -- prefetch.global.L2 -- was measured: no gain, 674 vs 675 GCUPS)
	if (ra == R) cs[0] = seed;
	// diagonals [fa, fb] on which the block is strictly inside the band (ks_tile_step_fast): en0(r) >= t0 + 19 and st0(r) < t0
	int fa, fb;
	ks_fast_range(c, k, ra, rb, fa, fb);
	int r = ra;
	ks_u4 *pc = cs + (size_t)(ra - R + 1) * sst, *pb = best + (size_t)(ra - R) * sst;    // this diagonal's records of the block on the left
	while (r <= rb) {
		ks_u4 co, bo;
		if (r == fa) {
			int st0 = ks_imax(ks_imax(0, r - c.qlen + 1), (r - c.w + 1) >> 1);
#if defined(__CUDA_ARCH__) && defined(KS_UNROLL)
#define KS_PRAGMA_(x) _Pragma(#x)
#define KS_PRAGMA_UNROLL(n) KS_PRAGMA_(unroll n)
			KS_PRAGMA_UNROLL(KS_UNROLL)
#endif
			for (; r <= fb; ++r, pc += sst, pb += sst) {
				const ks_u4 ccur = *pc, bin = KS_APX(CIG) ? ccur : *pb;          // (approximate max: no arg-max stream)
				const bool stop = ks_tile_step_fast<KIND, CIG>(P, c, ez, T, r, st0, cprev, ccur.x, bin, co, bo, prow, ez.apx_H0, ez.apx_t, ez.apx_r);
				*pc = co; if (!KS_APX(CIG)) *pb = bo;
				cprev = ccur;
				if (stop) { done = true; return; }
				st0 = ks_imax(ks_imax(0, r - c.qlen + 2), (r - c.w + 2) >> 1);
			}
			continue;
		}
		const ks_u4 ccur = *pc, bin = KS_APX(CIG) ? ccur : *pb;
		const bool stop = ks_tile_step<KIND, CIG>(P, c, ez, T, r, cprev, ccur, bin, save_left, co, bo, prow, ez.apx_H0, ez.apx_t, ez.apx_r);
		*pc = co; if (!KS_APX(CIG)) *pb = bo;
		cprev = ccur;
		if (stop) { done = true; return; }
		++r; pc += sst; pb += sst;
	}
	ks_tile_end<KIND>(P, c, T, save);
}
