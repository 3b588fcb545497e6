// ksw2_b200.cu -- kernels + C-ABI host side of the B200-native ksw2 hot path (sm_100a).
//
// Kernels (template arguments: KIND = extz2 / extd2 / exts2 semantics; CIG = 0 score only, 1 / 2 direction bytes for left- / right-aligned gaps,
//          + 4 = the approximate-max variants, KSW_EZ_APPROX_MAX)
//   ks_fill_kernel<KIND,CIG>       persistent grid; each warp pulls 32 jobs at a time from an atomic counter; one THREAD runs one
//                                  alignment with the register-resident tile engine (ksw2_tile.cuh): panels of diagonals, blocks left
//                                  to right.  Per-thread carry/arg-max streams live interleaved in shared memory, per-block saved
//                                  state in an L2-resident scratch arena, direction bytes stream to HBM ([pair][block][row][16]).
//   ks_fill_warp_kernel<..>        the same tiles with one WARP per alignment (diagonal-skewed wavefront over the lanes, 32 blocks per
//                                  wave): few pairs with a wide band
//   ks_fill_ring_kernel<..>        one WARP per alignment on the ring schedule (block k on lane k & 31 for its whole life): banded pairs
//                                  that are too few for a thread each -- long CIGAR pairs
//   ks_fill_cta_kernel<..>         the wavefront over the 256 threads of a CTA: a handful of very long pairs (latency)
//   ks_traceback_kernel            one thread per job: the ksw_backtrack state machine (ksw2.h:129-161) over the direction
//                                  bytes, two passes (count, then write run-length ops into a compacted CIGAR buffer).
//   ks_encode_kernel / ks_jobs_uniform_kernel   pre-coded sequences; job table of equal-length batches written on the device
//   ks_rows_kernel, ks_gg2_kernel, ks_extf2_kernel (+ tracebacks)   the other ksw2.h entry points: simple one-thread-per-pair kernels
//                                  (ksw2_rows.cuh, ksw2_gg2.cuh, ksw2_extf2.cuh); ks_scalar_kernel: the round-1 approximate-max kernel, kept
//                                  as a second opinion (KSW2B_SCALAR_APPROX=1)
// Host side: contexts, plans (job table with a band per pair, chunking of the direction arena, choice of the schedule per chunk), the
// pipelined batch call ksw2b_align_ex, device sets (ksw2b_multi_*), the drop-in single-pair entry points with their call-combining layer.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
#include "ksw2_pair.cuh"
#include "ksw2_params.h"
#include "ksw2_scalar.cuh"
#include "ksw2_rows.cuh"
#include "ksw2_extf2.cuh"
#include "ksw2_gg2.cuh"
#include "../../include/ksw2_b200.h"

// ------------------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------------------
struct KsJob { int64_t qoff, toff, poff, teoff, qeoff, soff; int32_t qlen, tlen, idx, w; };   // poff: direction arena offset (16-byte words); teoff/qeoff: byte offsets into the coded-sequence arenas; w: this pair's effective band (resolved on the host: ks_eff_w)

// One warp per pair: writes the coded target (block words in register lane order) and the coded, reversed, padded query.
__global__ void ks_encode_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs,
                                 const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, uint8_t *tenc, uint8_t *qenc)
{
	const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (w >= njobs) return;
	const KsJob job = jobs[w];
	if (job.qlen <= 0 || job.tlen <= 0) return;
	const uint8_t *t = tcat + job.toff, *q = qcat + job.qoff;
	uint8_t *te = tenc + job.teoff, *qe = qenc + job.qeoff;
	const int nt = ((job.tlen + 15) >> 4) << 4, nq = (int)ks_qenc_bytes(job.qlen);
	for (int i = lane; i < nt; i += 32) te[(i & ~15) + ks_perm_pos(i & 15)] = ks_enc_t(P, t, job.tlen, i);
	for (int i = lane; i < nq; i += 32) qe[i] = ks_enc_q(P, q, job.qlen, i - KS_QPADL);
}

// Register budgets.  The single-gap kernels are tuned for four 96-thread CTAs per SM (168 registers); the dual-gap / splice kernels hold two or
// three more state arrays and run three 96-thread CTAs per SM at ~200 registers.  Asked for three resident CTAs (or left alone) ptxas squeezes
// them into 168 registers with spills, which is slower (round 1: -4.6 %, round 2: -17 % on the 5 kb CIGAR workload); asked for two it takes
// the registers it needs (198 - 232) and three 96-thread CTAs still fit up to 227.
#ifdef KS_LB_B
#define KS_LB __launch_bounds__(KS_LB_T, KS_LB_B)        // experiments: trade registers for resident CTAs
#else
#define KS_LB __launch_bounds__(KIND == KS_Z ? 128 : 96, KIND == KS_Z ? 3 : 2)
#endif
#define KS_MAX_TPB(KIND) ((KIND) == KS_Z ? 128 : 96)
template<int KIND, int CIG>
__global__ void KS_LB
ks_fill_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs, unsigned long long *counter,
               const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const uint8_t *__restrict__ jcat,
               const uint8_t *__restrict__ tenc, const uint8_t *__restrict__ qenc, ks_u4 *save_arena, size_t save_stride, ks_u4 *parena, KsResult *res, int C)
{
	extern __shared__ uint4 ks_smem[];
	const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31;
	ks_u4 *cs = ks_smem + tid, *best = cs + (size_t)(C + 1) * nthr;
	ks_u4 *save = save_arena + ((size_t)blockIdx.x * nthr + tid) * save_stride;
	for (;;) {
		unsigned long long g = 0;
		if (lane == 0) g = atomicAdd(counter, 1ULL);
		g = __shfl_sync(0xffffffffu, g, 0);
		if ((long long)(g * 32ULL) >= njobs) break;
		const long long j = (long long)(g * 32ULL) + lane;
		if (j < njobs) {
			const KsJob job = jobs[j];
			KsResult out;
			KsEz ez; ks_ez_reset(ez);
			KsPair c;
			c.query = qcat + job.qoff; c.target = tcat + job.toff; c.junc = jcat ? jcat + job.toff : (const uint8_t*)0;
			c.tenc = tenc + job.teoff; c.qenc = qenc + job.qeoff + KS_QPADL;
			c.qlen = job.qlen; c.tlen = job.tlen;
			c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
			if (c.qlen > 0 && c.tlen > 0) {
				const int prows = ks_prows(c.qlen, c.tlen, c.w);
				ks_pair_fill<KIND, CIG>(P, c, ez, C, save, cs, best, nthr, KS_DIR(CIG) ? parena + job.poff : (ks_u4*)0, prows);
				ks_store_result(ez, out);
				ks_pick_start(P, c, ez, out);
			} else { ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; }
			res[job.idx] = out;
		}
		__syncwarp();
	}
}

// Warp-cooperative fill (ksw2_pair.cuh: ks_pair_fill_warp): one WARP per alignment, for batches with too few (long) pairs to
// fill the GPU with one thread each.  Per warp in shared memory: record ring (256 words), window of the incoming stream (66 words), ez
// scalars; the two inter-wave streams (4 * (C + 1) words per warp) live in global memory (wv_arena).
#define KS_WARP_SMEM_WORDS (256 + 66 + 4)
#define KS_WARP_WV_WORDS(C) (4 * ((size_t)(C) + 1))
template<int KIND, int CIG>
__global__ void __launch_bounds__(128)
ks_fill_warp_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs, unsigned long long *counter,
                    const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const uint8_t *__restrict__ jcat,
                    const uint8_t *__restrict__ tenc, const uint8_t *__restrict__ qenc, ks_u4 *save_arena, size_t save_stride, ks_u4 *wv_arena, ks_u4 *parena, KsResult *res, int C)
{
	__shared__ uint4 ks_wsm[4 * KS_WARP_SMEM_WORDS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
	ks_u4 *base = ks_wsm + (size_t)warp * KS_WARP_SMEM_WORDS;
	ks_u4 *ring = base, *inw = base + 256;
	KsWarpShared *ezs = (KsWarpShared*)(base + 256 + 66);
	ks_u4 *save = save_arena + ((size_t)blockIdx.x * wpc + warp) * save_stride;
	ks_u4 *wv = wv_arena + ((size_t)blockIdx.x * wpc + warp) * KS_WARP_WV_WORDS(C);
	for (;;) {
		unsigned long long g = 0;
		if (lane == 0) g = atomicAdd(counter, 1ULL);
		g = __shfl_sync(0xffffffffu, g, 0);
		if ((long long)g >= njobs) break;
		const KsJob job = jobs[g];
		KsPair c;
		c.query = qcat + job.qoff; c.target = tcat + job.toff; c.junc = jcat ? jcat + job.toff : (const uint8_t*)0;
		c.tenc = tenc + job.teoff; c.qenc = qenc + job.qeoff + KS_QPADL;
		c.qlen = job.qlen; c.tlen = job.tlen;
		c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
		if (c.qlen > 0 && c.tlen > 0) {
			ks_pair_fill_warp<KIND, CIG, 32>(P, c, ezs, C, save, ring, inw, wv, KS_DIR(CIG) ? parena + job.poff : (ks_u4*)0, ks_prows(c.qlen, c.tlen, c.w));
			__syncwarp();
			if (lane == 0) { KsResult out; ks_store_result(ezs->ez, out); ks_pick_start(P, c, ezs->ez, out); res[job.idx] = out; }
		} else if (lane == 0) { KsResult out; KsEz ez; ks_ez_reset(ez); ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; res[job.idx] = out; }
		__syncwarp();
	}
}

// Ring schedule of the warp-cooperative fill (ksw2_pair.cuh: ks_pair_fill_ring): one WARP per alignment, BANDED pairs (effective band <= 512):
// block k stays on lane k & 31 for its whole life, no panels, no saved state, ~31 of 32 lanes busy.  What long CIGAR pairs run on: their
// direction bytes (MBs per pair) bound the pairs in flight, and a warp per pair needs ~50 x fewer pairs in flight than a thread per pair.
#define KS_RING_SMEM_WORDS (256 + 4)
template<int KIND, int CIG>
__global__ void __launch_bounds__(128)
ks_fill_ring_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs, unsigned long long *counter,
                    const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const uint8_t *__restrict__ jcat,
                    const uint8_t *__restrict__ tenc, const uint8_t *__restrict__ qenc, ks_u4 *save_arena, size_t save_stride, ks_u4 *parena, KsResult *res)
{
	__shared__ uint4 ks_rsm[4 * KS_RING_SMEM_WORDS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
	ks_u4 *ring = ks_rsm + (size_t)warp * KS_RING_SMEM_WORDS;
	KsWarpShared *ezs = (KsWarpShared*)(ring + 256);
	ks_u4 *save = save_arena + ((size_t)blockIdx.x * wpc + warp) * save_stride;
	for (;;) {
		unsigned long long g = 0;
		if (lane == 0) g = atomicAdd(counter, 1ULL);
		g = __shfl_sync(0xffffffffu, g, 0);
		if ((long long)g >= njobs) break;
		const KsJob job = jobs[g];
		KsPair c;
		c.query = qcat + job.qoff; c.target = tcat + job.toff; c.junc = jcat ? jcat + job.toff : (const uint8_t*)0;
		c.tenc = tenc + job.teoff; c.qenc = qenc + job.qeoff + KS_QPADL;
		c.qlen = job.qlen; c.tlen = job.tlen;
		c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
		if (c.qlen > 0 && c.tlen > 0) {
			ks_pair_fill_ring<KIND, CIG>(P, c, ezs, save, ring, KS_DIR(CIG) ? parena + job.poff : (ks_u4*)0, ks_prows(c.qlen, c.tlen, c.w));
			__syncwarp();
			if (lane == 0) { KsResult out; ks_store_result(ezs->ez, out); ks_pick_start(P, c, ezs->ez, out); res[job.idx] = out; }
		} else if (lane == 0) { KsResult out; KsEz ez; ks_ez_reset(ez); ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; res[job.idx] = out; }
		__syncwarp();
	}
}

// CTA-cooperative fill: the same wavefront over all KS_CTA_LANES threads of a CTA, one alignment per CTA (ks_pair_fill_warp with NL lanes,
// __syncthreads per step): for launches of very few very long pairs, where the latency of a pair is what counts.
#define KS_CTA_LANES 256
#define KS_CTA_SMEM_WORDS (8 * KS_CTA_LANES + 66 + 4)
template<int KIND, int CIG>
__global__ void __launch_bounds__(KS_CTA_LANES)
ks_fill_cta_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs, unsigned long long *counter,
                   const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const uint8_t *__restrict__ jcat,
                   const uint8_t *__restrict__ tenc, const uint8_t *__restrict__ qenc, ks_u4 *save_arena, size_t save_stride, ks_u4 *wv_arena, ks_u4 *parena, KsResult *res, int C)
{
	__shared__ uint4 ks_csm[KS_CTA_SMEM_WORDS];
	__shared__ unsigned long long ks_next;
	ks_u4 *ring = ks_csm, *inw = ks_csm + 8 * KS_CTA_LANES;
	KsWarpShared *ezs = (KsWarpShared*)(ks_csm + 8 * KS_CTA_LANES + 66);
	ks_u4 *save = save_arena + (size_t)blockIdx.x * save_stride;
	ks_u4 *wv = wv_arena + (size_t)blockIdx.x * KS_WARP_WV_WORDS(C);
	for (;;) {
		if (threadIdx.x == 0) ks_next = atomicAdd(counter, 1ULL);
		__syncthreads();
		const unsigned long long g = ks_next;
		__syncthreads();
		if ((long long)g >= njobs) break;
		const KsJob job = jobs[g];
		KsPair c;
		c.query = qcat + job.qoff; c.target = tcat + job.toff; c.junc = jcat ? jcat + job.toff : (const uint8_t*)0;
		c.tenc = tenc + job.teoff; c.qenc = qenc + job.qeoff + KS_QPADL;
		c.qlen = job.qlen; c.tlen = job.tlen;
		c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
		if (c.qlen > 0 && c.tlen > 0) {
			ks_pair_fill_warp<KIND, CIG, KS_CTA_LANES>(P, c, ezs, C, save, ring, inw, wv, KS_DIR(CIG) ? parena + job.poff : (ks_u4*)0, ks_prows(c.qlen, c.tlen, c.w));
			__syncthreads();
			if (threadIdx.x == 0) { KsResult out; ks_store_result(ezs->ez, out); ks_pick_start(P, c, ezs->ez, out); res[job.idx] = out; }
		} else if (threadIdx.x == 0) { KsResult out; KsEz ez; ks_ez_reset(ez); ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; res[job.idx] = out; }
		__syncthreads();
	}
}

// approximate-max mode (KSW_EZ_APPROX_MAX) as an in-order scalar sweep, one thread per job (ksw2_scalar.cuh): only with KSW2B_SCALAR_APPROX=1 -- the
// mode normally runs on the tile engine like everything else (ks_apx_step in ksw2_tile.cuh)
__global__ void ks_scalar_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs,
                                 const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const uint8_t *__restrict__ jcat,
                                 int8_t *scratch, ks_u4 *parena, KsResult *res)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	KsResult out; KsEz ez; ks_ez_reset(ez);
	KsPair c;
	c.query = qcat + job.qoff; c.target = tcat + job.toff; c.junc = jcat ? jcat + job.toff : (const uint8_t*)0; c.tenc = c.qenc = 0;
	c.qlen = job.qlen; c.tlen = job.tlen;
	c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
	if (c.qlen > 0 && c.tlen > 0) {
		ks_pair_scalar(P, c, ez, scratch + job.soff, (uint8_t*)(parena + job.poff), ks_prows(c.qlen, c.tlen, c.w));
		ks_store_result(ez, out); ks_pick_start(P, c, ez, out);
	} else { ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; }
	res[job.idx] = out;
}

__global__ void ks_traceback_kernel(const __grid_constant__ KsParams P, const KsJob *__restrict__ jobs, long long njobs,
                                    const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, const ks_u4 *parena, KsResult *res, uint32_t *cig, unsigned long long *cursor, long long cap)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	KsResult r = res[job.idx];
	if (r.tb_i < 0) return;
	KsPair c;
	c.junc = c.tenc = c.qenc = 0; c.qlen = job.qlen; c.tlen = job.tlen;
	c.query = qcat + job.qoff; c.target = tcat + job.toff;          // only read for KSW_EZ_EQX
	c.w = job.w; c.ndiag = c.qlen + c.tlen - 1; c.tlen_ = (c.tlen + 15) >> 4;
	const int prows = ks_prows(c.qlen, c.tlen, c.w);
	const uint8_t *pb = (const uint8_t*)(parena + job.poff);
	const int n = ks_traceback(P, c, pb, prows, r.tb_i, r.tb_j, 0, 0);
	const unsigned long long off = atomicAdd(cursor, (unsigned long long)n);
	if ((long long)(off + n) <= cap) ks_traceback(P, c, pb, prows, r.tb_i, r.tb_j, cig + off, n);
	res[job.idx].n_cigar = n; res[job.idx].cigar_off = (int64_t)off;
}

// Row-wise entry points ksw_extz / ksw_extd (ksw2_rows.cuh): persistent grid, each warp pulls 32 jobs at a time; one thread
// per pair.  eh rows of the warp's 32 pairs are interleaved word by word in the warp's scratch slot (every access = one line).
__global__ void __launch_bounds__(128)
ks_rows_kernel(const __grid_constant__ KsRowsParams RP, const KsJob *__restrict__ jobs, long long njobs, unsigned long long *counter,
               const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, int32_t *scratch, size_t warp_words, ks_u4 *parena, KsResult *res)
{
	const int lane = threadIdx.x & 31;
	int32_t *eh = scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * warp_words + lane;
	for (;;) {
		unsigned long long g = 0;
		if (lane == 0) g = atomicAdd(counter, 1ULL);
		g = __shfl_sync(0xffffffffu, g, 0);
		if ((long long)(g * 32ULL) >= njobs) break;
		const long long j = (long long)(g * 32ULL) + lane;
		if (j < njobs) {
			const KsJob job = jobs[j];
			KsResult out; KsEz ez; ks_ez_reset(ez);
			if (job.qlen > 0 && job.tlen > 0) {
				ks_rows_fill(RP, qcat + job.qoff, job.qlen, tcat + job.toff, job.tlen, eh, 32, (uint8_t*)(parena + job.poff), ez);
				ks_store_result(ez, out);
				ks_rows_pick_start(RP, job.qlen, job.tlen, ez, out);
			} else { ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0; }
			res[job.idx] = out;
		}
		__syncwarp();
	}
}

__global__ void ks_rows_traceback_kernel(const __grid_constant__ KsRowsParams RP, const KsJob *__restrict__ jobs, long long njobs,
                                         const ks_u4 *parena, KsResult *res, uint32_t *cig, unsigned long long *cursor, long long cap)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	const KsResult r = res[job.idx];
	if (r.tb_i < 0) return;
	const uint8_t *z = (const uint8_t*)(parena + job.poff);
	const int n = ks_rows_traceback(RP, job.qlen, job.tlen, z, r.tb_i, r.tb_j, 0, 0);
	const unsigned long long off = atomicAdd(cursor, (unsigned long long)n);
	if ((long long)(off + n) <= cap) ks_rows_traceback(RP, job.qlen, job.tlen, z, r.tb_i, r.tb_j, cig + off, n);
	res[job.idx].n_cigar = n; res[job.idx].cigar_off = (int64_t)off;
}

// ksw_extf2_sse (ksw2_extf2.cuh): one thread per job, in-order; scratch = the reference's flat layout per pair
__global__ void ks_extf2_kernel(const __grid_constant__ KsExtfParams FP, const KsJob *__restrict__ jobs, long long njobs,
                                const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, uint8_t *scratch, KsResult *res)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	KsResult out; KsEz ez; ks_ez_reset(ez);
	if (job.qlen > 0 && job.tlen > 0) ks_extf2(FP, qcat + job.qoff, job.qlen, tcat + job.toff, job.tlen, scratch + job.soff, ez);
	ks_store_result(ez, out); out.tb_i = out.tb_j = -1; out.reach_end = 0;
	res[job.idx] = out;
}

// ksw_gg2 / ksw_gg2_sse (ksw2_gg2.cuh): one thread per job, in-order fill, then the rotated traceback
__global__ void ks_gg2_kernel(const __grid_constant__ KsGg2Params GP, const KsJob *__restrict__ jobs, long long njobs,
                              const uint8_t *__restrict__ qcat, const uint8_t *__restrict__ tcat, int8_t *scratch, ks_u4 *parena, KsResult *res, int with_cigar)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	KsResult out; KsEz ez; ks_ez_reset(ez);
	if (job.qlen > 0 && job.tlen > 0) {
		ez.score = ks_gg2_fill(GP, qcat + job.qoff, job.qlen, tcat + job.toff, job.tlen, scratch + job.soff, with_cigar ? (uint8_t*)(parena + job.poff) : (uint8_t*)0);
		ez.n_diag = job.qlen + job.tlen - 1;
	}
	ks_store_result(ez, out); out.reach_end = 0;
	out.tb_i = (with_cigar && job.qlen > 0 && job.tlen > 0) ? job.tlen - 1 : -1; out.tb_j = out.tb_i < 0 ? -1 : job.qlen - 1;
	res[job.idx] = out;
}
__global__ void ks_gg2_traceback_kernel(const __grid_constant__ KsGg2Params GP, const KsJob *__restrict__ jobs, long long njobs,
                                        const ks_u4 *parena, KsResult *res, uint32_t *cig, unsigned long long *cursor, long long cap)
{
	const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const KsJob job = jobs[j];
	if (res[job.idx].tb_i < 0) return;
	const uint8_t *dir = (const uint8_t*)(parena + job.poff);
	const int n = ks_gg2_traceback(GP, job.qlen, job.tlen, dir, 0, 0);
	const unsigned long long off = atomicAdd(cursor, (unsigned long long)n);
	if ((long long)(off + n) <= cap) ks_gg2_traceback(GP, job.qlen, job.tlen, dir, cig + off, n);
	res[job.idx].n_cigar = n; res[job.idx].cigar_off = (int64_t)off;
}

// Job table of a batch whose pairs all have the same lengths (score-only runs): every field is a closed form of the index, so the
// table is written on the device instead of being built on the host (1.6 ms per million pairs) and uploaded (64 B per pair).
__global__ void ks_jobs_uniform_kernel(KsJob *jobs, long long lo, long long hi, long long q0, long long t0, int qlen, int tlen, int w,
                                       long long te_stride, long long qe_stride, long long s_stride)
{
	const long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	KsJob j;
	j.qoff = q0 + i * qlen; j.toff = t0 + i * tlen; j.poff = 0; j.teoff = i * te_stride; j.qeoff = i * qe_stride; j.soff = i * s_stride;
	j.qlen = qlen; j.tlen = tlen; j.idx = (int32_t)i; j.w = w;
	jobs[i] = j;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int ks_fail(int code, const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
	return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ks_fail(-10, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct DevBuf {
	void *p = 0; size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = 0; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { p = 0; return -1; } want = bytes; }
		cap = want; return 0;
	}
	void release() { if (p) cudaFree(p); p = 0; cap = 0; }
};
struct PinBuf {
	void *p = 0; size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return 0;
		if (p) cudaFreeHost(p);
		p = 0; cap = 0;
		if (cudaMallocHost(&p, bytes + bytes / 8 + 256) != cudaSuccess) { p = 0; return -1; }
		cap = bytes + bytes / 8 + 256; return 0;
	}
	void release() { if (p) cudaFreeHost(p); p = 0; cap = 0; }
};

struct ksw2b_plan;
struct ksw2b_ctx {
	int device = 0, num_sm = 0, num_dev = 1;
	int panel = 15, threads = 96, ctas_per_sm = 4;    // measured best on the 150 bp workload (profiles/r1_tuning.txt)
	bool auto_panel = true;                           // taller panels for launches that under-fill the GPU; off once a caller sets a panel
	int mode = 0, wpanel = 1024;                      // 0 auto, 1 one thread per pair, 2 one warp per pair; panel height of the warp mode (its streams live in global memory)
	size_t smem_optin = 0, smem_sm = 0;
	DevBuf d_q, d_t, d_j, d_jobs, d_res, d_save, d_wv, d_parena, d_cig, d_ctr, d_mat, d_tenc, d_qenc, d_scal;
	PinBuf h_jobs, h_res, h_q, h_t, h_j;   // h_q/h_t/h_j: staging of the array-of-pointers batch calls
	std::vector<uint32_t> cig_host;     // concatenated CIGARs of the last fetch
	cudaStream_t s_in = 0, s_job = 0, s_cmp = 0, s_cmp2 = 0, s_out = 0;
	std::vector<cudaEvent_t> ev;
	unsigned long long last_h2d = 0, last_d2h = 0;      // bytes the last ksw2b_align moved over PCIe (inputs + job table; results + CIGARs)
	bool scalar_approx = false;         // KSW2B_SCALAR_APPROX=1
	int l2_persist = 0;                 // KSW2B_L2PERSIST=1: the save area of the thread-per-pair kernels is marked persisting in L2 (access policy window)
	size_t l2_window_max = 0, l2_persist_max = 0;
	const void *l2_win_ptr[2] = {0, 0}; size_t l2_win_bytes[2] = {0, 0}; cudaStream_t l2_win_stream[2] = {0, 0};   // window currently set on s_cmp / s_cmp2 (or the caller's stream)
	bool timing = false;                // ksw2b_set_timing: ksw2b_align brackets its kernels with CUDA events
	cudaEvent_t tm[3] = {0, 0, 0};      // first kernel of the call; end of the work on each compute stream
	double last_fill_ms = 0, last_span_ms = 0; int last_fill_launches = 0, last_launches = 0;
	ksw2b_plan *live_plan = 0;          // plans borrow the buffers above: one live plan per context (plan_build refuses a second one)
	std::unordered_map<const void*, cudaFuncAttributes> fattr;   // kernel attributes, asked once per kernel
};
// blocking upload of a small table that kernels on the context's NON-BLOCKING streams will read: cudaMemcpy from pageable memory may
// return before the DMA has landed and those streams do not order against the legacy stream, so wait for it explicitly
static int upload_small(DevBuf &b, const void *src, size_t bytes)
{
	if (b.ensure(bytes) || cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess || cudaStreamSynchronize(0) != cudaSuccess) { cudaGetLastError(); return -1; }
	return 0;
}

struct Chunk { int64_t lo, hi; int64_t pwords, cigcap; int seg; bool warp, cta, ring; int max_tlen_; bool force_ring; };   // warp: this chunk runs one WARP per pair (ring: on the ring schedule); cta: one CTA per pair
struct Seg { int64_t lo, hi; size_t c0, c1; };

struct ksw2b_plan {
	ksw2b_ctx *ctx = 0;
	KsParams P;
	int prep = KS_PREP_OK;
	int cig = 0;                       // 0 score only, 1 left, 2 right
	int64_t n = 0, cells = -1;
	int launches = 0;
	int max_tlen_ = 1;
	KsJob *jobs = 0;                   // pinned (ctx->h_jobs); sorted inside each segment
	std::vector<Chunk> chunks;
	std::vector<Seg> segs;
	size_t save_words = 0;             // 16-byte words of ONE save arena; the context holds two (launches on alternating streams)
	size_t wv_words = 0; int wpanel = 0; // warp mode: words of ONE inter-wave stream arena, panel height
	int slot = 0;                      // which of the two the next launch uses
	int grid = 0, grid_warp = 0;       // persistent grids of the thread-per-pair / warp-per-pair launches
	size_t save_stride_thread = 0, save_stride_warp = 0;
	int64_t tenc_bytes = 0, qenc_bytes = 0, scal_bytes = 0;
	bool approx = false, warp_mode = false;
	bool rows = false;                 // ksw_extz / ksw_extd: the row-wise kernels (ksw2_rows.cuh)
	KsRowsParams RP;
	bool uniform = false;              // all pairs have the same lengths and no CIGAR is wanted: the job table is generated on the device
	int u_qlen = 0, u_tlen = 0, u_w = 0; int64_t u_q0 = 0, u_t0 = 0, u_te = 0, u_qe = 0, u_sc = 0;
	bool extf = false;                 // ksw_extf2_sse (ksw2_extf2.cuh)
	KsExtfParams FP;
	bool gg2 = false;                  // ksw_gg2 / ksw_gg2_sse (ksw2_gg2.cuh)
	KsGg2Params GP;
	int max_qlen = 1, rows_grid = 0;
	size_t rows_warp_words = 0;
	std::vector<int64_t> chunk_cig_used;
	int64_t cig_half = 0;              // CIGAR runs with several chunks: the staging buffer has two halves of this many words (chunk ci writes half ci & 1)
	cudaEvent_t cig_ev[2] = {0, 0};    // traceback of the chunk that wrote the half has finished
	int64_t drained = 0;               // chunks of the current run whose CIGAR words are on the host
	bool ran = false;
	bool timing = false;               // record CUDA events around every fill launch (ksw2b_plan_set_timing)
	std::vector<cudaEvent_t> tev;      // pairs (start, stop), one pair per fill launch of the last run
	size_t tev_used = 0;
	~ksw2b_plan() { for (auto e : tev) cudaEventDestroy(e); for (auto e : cig_ev) if (e) cudaEventDestroy(e); if (ctx && ctx->live_plan == this) ctx->live_plan = 0; }
};

extern "C" const char *ksw2b_last_error(void) { return g_err; }

extern "C" ksw2b_ctx_t *ksw2b_create(int device)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ks_fail(-1, "no CUDA device: ksw2_b200 has no CPU path"); return 0; }
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
	if (device >= ndev) { ks_fail(-1, "device %d out of range (%d devices)", device, ndev); return 0; }
	if (cudaSetDevice(device) != cudaSuccess) { ks_fail(-1, "cudaSetDevice(%d) failed", device); return 0; }
	cudaDeviceProp pr;
	if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { ks_fail(-1, "cudaGetDeviceProperties failed"); return 0; }
	ksw2b_ctx *c = new ksw2b_ctx();
	// knobs for experiments (scripts/): KSW2B_PANEL / KSW2B_THREADS / KSW2B_CTAS as ksw2b_set_tuning, KSW2B_MODE / KSW2B_WPANEL as ksw2b_set_mode
	{ const char *e;
	  if ((e = getenv("KSW2B_PANEL")) && atoi(e) > 0) { c->panel = atoi(e); c->auto_panel = false; }
	  if ((e = getenv("KSW2B_THREADS")) && atoi(e) > 0) c->threads = atoi(e) > 128 ? 128 : (atoi(e) + 31) / 32 * 32;
	  if ((e = getenv("KSW2B_CTAS")) && atoi(e) > 0) c->ctas_per_sm = atoi(e);
	  if ((e = getenv("KSW2B_MODE")) && atoi(e) >= 0 && atoi(e) <= 4) c->mode = atoi(e);
	  if ((e = getenv("KSW2B_WPANEL")) && atoi(e) > 0) c->wpanel = atoi(e);
	  if ((e = getenv("KSW2B_SCALAR_APPROX")) && atoi(e) > 0) c->scalar_approx = true; }
	c->l2_window_max = (size_t)pr.accessPolicyMaxWindowSize; c->l2_persist_max = (size_t)pr.persistingL2CacheMaxSize;
	{ const char *e = getenv("KSW2B_L2PERSIST"); c->l2_persist = e ? atoi(e) : 0; }
	if (c->l2_persist && c->l2_persist_max > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_max);
	c->device = device; c->num_sm = pr.multiProcessorCount; c->num_dev = ndev; c->smem_optin = pr.sharedMemPerBlockOptin; c->smem_sm = pr.sharedMemPerMultiprocessor > 1024 ? pr.sharedMemPerMultiprocessor - 1024 : pr.sharedMemPerMultiprocessor;
	return c;
}

extern "C" void ksw2b_destroy(ksw2b_ctx_t *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	c->d_q.release(); c->d_t.release(); c->d_j.release(); c->d_jobs.release(); c->d_res.release(); c->d_save.release(); c->d_wv.release();
	c->d_parena.release(); c->d_cig.release(); c->d_ctr.release(); c->d_mat.release(); c->d_tenc.release(); c->d_qenc.release(); c->d_scal.release();
	c->h_jobs.release(); c->h_res.release(); c->h_q.release(); c->h_t.release(); c->h_j.release();
	if (c->s_in) cudaStreamDestroy(c->s_in);
	if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
	if (c->s_cmp2) cudaStreamDestroy(c->s_cmp2);
	if (c->s_job) cudaStreamDestroy(c->s_job);
	if (c->s_out) cudaStreamDestroy(c->s_out);
	for (auto e : c->ev) cudaEventDestroy(e);
	for (auto e : c->tm) if (e) cudaEventDestroy(e);
	delete c;
}

extern "C" void ksw2b_set_tuning(ksw2b_ctx_t *c, int panel, int threads, int ctas_per_sm)
{
	if (!c) return;
	if (panel > 0) { c->panel = panel; c->auto_panel = false; }
	if (threads > 0) c->threads = threads > 128 ? 128 : (threads + 31) / 32 * 32;
	if (ctas_per_sm > 0) c->ctas_per_sm = ctas_per_sm;
}

extern "C" void ksw2b_set_mode(ksw2b_ctx_t *c, int mode, int warp_panel)
{
	if (!c) return;
	if (mode >= 0 && mode <= 4) c->mode = mode;
	if (warp_panel > 0) c->wpanel = warp_panel;
}

extern "C" void ksw2b_set_timing(ksw2b_ctx_t *c, int on) { if (c) c->timing = on != 0; }
extern "C" void ksw2b_last_timing(ksw2b_ctx_t *c, double *fill_ms, int *fill_launches, double *span_ms, int *launches)
{
	if (fill_ms) *fill_ms = c ? c->last_fill_ms : 0;
	if (fill_launches) *fill_launches = c ? c->last_fill_launches : 0;
	if (span_ms) *span_ms = c ? c->last_span_ms : 0;
	if (launches) *launches = c ? c->last_launches : 0;
}

extern "C" void ksw2b_last_transfer_bytes(ksw2b_ctx_t *c, unsigned long long *h2d, unsigned long long *d2h)
{
	if (h2d) *h2d = c ? c->last_h2d : 0;
	if (d2h) *d2h = c ? c->last_d2h : 0;
}

extern "C" void *ksw2b_host_alloc(size_t bytes) { void *p = 0; if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return 0; } return p; }
extern "C" void ksw2b_host_free(void *p) { if (p) cudaFreeHost(p); }

// effective band of one pair (ksw2_extz2_sse.c:64: w < 0 or wider than the longer sequence = no band); exts2 has none (ksw2_exts2_sse.c:179-182)
static inline int ks_eff_w(int kind, int w, int qlen, int tlen)
{
	const int mx = qlen > tlen ? qlen : tlen;
	return (kind == KS_S || w < 0 || w > mx) ? mx : w;
}

// work estimate of one pair whose band is already resolved (KsJob::w): lanes of the band once it is full x diagonals / 2 (+ the traceback's share)
static inline int64_t pair_cost_w(int we, int ql, int tl, bool cigar)
{
	if (ql <= 0 || tl <= 0) return 1;
	const int shortest = ql < tl ? ql : tl;
	const int64_t band = std::min<int64_t>(shortest, 2ll * we + 1);
	return band * ((int64_t)ql + tl) / 2 + 1 + (cigar ? ql + tl : 0);
}

static int64_t band_cells(int qlen, int tlen, int w)
{
	int64_t s = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0 = std::max(std::max(0, r - qlen + 1), (r - w + 1) >> 1), en0 = std::min(std::min(tlen - 1, r), (r + w) >> 1);
		if (st0 > en0) break;
		s += en0 - st0 + 1;
	}
	return s;
}

static int64_t rows_cells(int qlen, int tlen, int w)       // row-wise kernels: columns max(0,i-w) .. min(qlen-1,i+w) of every target row i
{
	int64_t s = 0;
	for (int i = 0; i < tlen; ++i) { const int st = i > w ? i - w : 0, en = i + w < qlen - 1 ? i + w : qlen - 1; if (en >= st) s += en - st + 1; }
	return s;
}

// device copy of the job records [lo, hi): generated in place for uniform batches, copied from the pinned host table otherwise
static int upload_jobs(ksw2b_plan *pl, int64_t lo, int64_t hi, cudaStream_t st)
{
	ksw2b_ctx *ctx = pl->ctx;
	if (hi <= lo) return 0;
	if (pl->uniform) {
		ks_jobs_uniform_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>((KsJob*)ctx->d_jobs.p, lo, hi, pl->u_q0, pl->u_t0, pl->u_qlen, pl->u_tlen, pl->u_w, pl->u_te, pl->u_qe, pl->u_sc);
		CK(cudaGetLastError());
		return 0;
	}
	CK(cudaMemcpyAsync((KsJob*)ctx->d_jobs.p + lo, pl->jobs + lo, sizeof(KsJob) * (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
	return 0;
}

// Builds the job table (nseg contiguous input segments, jobs sorted inside a segment so that the 32 jobs of a warp share a
// geometry where possible), cuts segments into chunks that fit the direction arena, sizes and allocates all device scratch.
static ksw2b_plan *plan_build(ksw2b_ctx *ctx, const ksw2b_params_t *par, int64_t n, const int64_t *qoff, const int64_t *toff, const int32_t *wv,
                              const std::vector<int64_t> &bounds, bool upload)
{
	if (!ctx || !par || n < 0) { ks_fail(-2, "bad arguments"); return 0; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { ks_fail(-1, "cudaSetDevice failed"); return 0; }
	// A plan borrows the context's buffers (job table, coded sequences, scratch, direction arena): ONE live plan per context
	if (ctx->live_plan) { ks_fail(-4, "context busy: another plan is alive on it (one live plan per context; destroy it first)"); return 0; }
	if (wv && par->kind > KSW2B_EXTS2) { ks_fail(-2, "a per-pair band is only supported for the ext*2 kinds"); return 0; }
	ksw2b_plan *pl = new ksw2b_plan();
	pl->ctx = ctx; pl->n = n; ctx->live_plan = pl;
	std::vector<int8_t> smat((size_t)std::max(1, par->m * par->m));
	pl->rows = par->kind == KSW2B_EXTZ || par->kind == KSW2B_EXTD || par->kind == KSW2B_GG;
	pl->extf = par->kind == KSW2B_EXTF2;
	pl->gg2 = par->kind == KSW2B_GG2 || par->kind == KSW2B_GG2_SSE;
	if (pl->gg2) {                                       // no early-outs in the reference (ksw2_gg2.c:4-24)
		if (par->m <= 0 || !par->mat) { ks_fail(-2, "ksw_gg2 needs a scoring matrix"); delete pl; return 0; }
		memset(&pl->P, 0, sizeof pl->P);
		pl->P.kind = par->kind; pl->P.flag = par->flag & KSF_SCORE_ONLY; pl->P.m = par->m; pl->P.w = par->w;
		pl->GP.sse = par->kind == KSW2B_GG2_SSE; pl->GP.m = par->m; pl->GP.q = (int8_t)par->q; pl->GP.e = (int8_t)par->e; pl->GP.w = par->w;
		if (upload_small(ctx->d_mat, par->mat, smat.size())) {
			ks_fail(-10, "matrix upload failed"); delete pl; return 0;
		}
		pl->GP.mat = (const int8_t*)ctx->d_mat.p;
		pl->prep = KS_PREP_OK;
	} else if (pl->extf) {                                      // no set-up section and no early-outs in the reference (ksw2_extf2_sse.c:11-24)
		memset(&pl->P, 0, sizeof pl->P);
		pl->P.kind = par->kind; pl->P.flag = KSF_SCORE_ONLY; pl->P.w = par->w;
		pl->FP.mch = (int8_t)par->q; pl->FP.mis = (int8_t)par->q2; pl->FP.e = (int8_t)par->e; pl->FP.w = par->w; pl->FP.xdrop = par->zdrop;
		pl->prep = KS_PREP_OK;
	} else if (pl->rows) {                                      // row-wise entry points: no set-up section, no early-outs in the reference (ksw2_extz.c:6-30)
		if (par->m <= 0 || !par->mat) { ks_fail(-2, "ksw_extz/ksw_extd need a scoring matrix"); delete pl; return 0; }
		memset(&pl->P, 0, sizeof pl->P);
		pl->P.kind = par->kind; pl->P.flag = par->flag; pl->P.m = par->m; pl->P.w = par->w;
		KsRowsParams &R = pl->RP;
		R.kind = par->kind == KSW2B_EXTZ ? KS_ROWZ : par->kind == KSW2B_GG ? KS_ROWG : KS_ROWD; R.m = par->m; R.gapo = (int8_t)par->q; R.gape = (int8_t)par->e;
		R.gapo2 = (int8_t)par->q2; R.gape2 = (int8_t)par->e2; R.w = par->w; R.zdrop = par->zdrop; R.flag = par->flag;
		if (upload_small(ctx->d_mat, par->mat, smat.size())) {
			ks_fail(-10, "matrix upload failed"); delete pl; return 0;
		}
		R.mat = (const int8_t*)ctx->d_mat.p;
		pl->prep = KS_PREP_OK;
	} else
	pl->prep = ks_prepare_params(pl->P, par->kind, par->m, par->mat, par->q, par->e, par->q2, par->e2, par->w, par->zdrop, par->end_bonus,
	                             par->flag, par->noncan, par->junc_bonus, smat.data(), 0);
	pl->cig = (pl->extf || (par->flag & KSF_SCORE_ONLY)) ? 0 : (!pl->gg2 && (par->flag & KSF_RIGHT)) ? 2 : 1;
	// KSW_EZ_APPROX_MAX runs on the tile engine (ks_apx_step); KSW2B_SCALAR_APPROX=1 selects the in-order scalar kernel instead (a second opinion for tests)
	pl->approx = !pl->rows && !pl->extf && !pl->gg2 && (par->flag & KSF_APPROX_MAX) != 0 && ctx->scalar_approx;
	if (!pl->rows && !pl->extf && !pl->gg2 && pl->prep == KS_PREP_OK && pl->P.smode == 1) {
		if (upload_small(ctx->d_mat, smat.data(), smat.size())) {
			ks_fail(-10, "matrix upload failed"); delete pl; return 0;
		}
		pl->P.mat = (const int8_t*)ctx->d_mat.p;
	}
	if (ctx->h_jobs.ensure(sizeof(KsJob) * (size_t)std::max<int64_t>(1, n))) { ks_fail(-11, "pinned job table allocation failed"); delete pl; return 0; }
	pl->jobs = (KsJob*)ctx->h_jobs.p;
	// direction arena budget: 88 % of what is free (+ what the arena already holds).  cudaMemGetInfo costs a fraction of a millisecond,
	// so it is only asked when a segment does not fit the arena the context already owns (small batches of the combining layer)
	int64_t arena_budget = -1;
	auto arena_words_max_for = [&](int64_t total_words) -> int64_t {
		if (const char *e = getenv("KSW2B_ARENA_MB")) { const int64_t mb = atoll(e); if (mb > 0) return mb * (1 << 20) / 16; }    // (tests: force several chunks on small batches)
		if (total_words * 16 <= (int64_t)ctx->d_parena.cap) return (int64_t)(ctx->d_parena.cap / 16);
		if (arena_budget < 0) {
			// what this plan still has to allocate besides the arena (upper bounds: thread-mode save area at the full grid, the CIGAR staging cap)
			const int SWx = pl->P.kind == KS_Z ? (int)KsSaveWords<KS_Z>::value : pl->P.kind == KS_D ? (int)KsSaveWords<KS_D>::value : (int)KsSaveWords<KS_S>::value;
			auto grow = [](size_t need, const DevBuf &b) -> size_t { return need > b.cap ? need + need / 8 + 256 : 0; };
			const size_t save_need = (pl->rows || pl->extf || pl->gg2) ? 0 : (size_t)(pl->cig ? 1 : 2) * ((size_t)ctx->num_sm * ctx->ctas_per_sm * ctx->threads + 32) * (size_t)pl->max_tlen_ * SWx * 16;
			const size_t others = grow(sizeof(KsJob) * (size_t)n, ctx->d_jobs) + grow(sizeof(KsResult) * (size_t)n, ctx->d_res) + grow(save_need, ctx->d_save) +
			                      grow((size_t)pl->tenc_bytes + 64, ctx->d_tenc) + grow((size_t)pl->qenc_bytes + 64, ctx->d_qenc) + grow((size_t)pl->scal_bytes + 64, ctx->d_scal) +
			                      grow((size_t)std::min<int64_t>(192ll << 20, qoff[n] + toff[n] + n) * 4, ctx->d_cig);
			size_t free_b = 0, tot_b = 0; cudaMemGetInfo(&free_b, &tot_b);
			const double avail = (double)free_b + (double)ctx->d_parena.cap - (double)others;
			arena_budget = (int64_t)(std::max(avail, 64.0 * 1024 * 1024) * 0.88 / 16.0);
		}
		return arena_budget;
	};
	const int64_t cig_words_max = 192ll << 20;           // 768 MiB of CIGAR words per chunk at most
	const int nseg = (int)bounds.size() - 1;
	bool all_uniform = false;
	// 1) job records (64 B each, memory bound): filled by several host threads; offsets into the coded-sequence arenas are
	//    assigned per thread range from a prefix over the ranges' byte totals
	{
		// host threads for the job table: at most 16, at most the host's cores divided by the GPUs of the box (one process per GPU is the usual
		// deployment: eight ranks that each spawn 16 threads on a 32-core host only get in each other's way), one per 64 k pairs
		const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, ctx->num_dev))), n / 65536));
		std::vector<int64_t> te(T + 1, 0), qe(T + 1, 0), sc(T + 1, 0); std::vector<int> mt(T, 1), mq(T, 1), uni(T, 1);
		const int q0len = n > 0 ? (int)(qoff[1] - qoff[0]) : 0, t0len = n > 0 ? (int)(toff[1] - toff[0]) : 0;
		const bool ok = pl->prep == KS_PREP_OK, approx = pl->approx || pl->extf || pl->gg2, extf = pl->extf, gg2 = pl->gg2;
		const int kind = pl->P.kind, w_all = pl->P.w;      // (exts2: P.w == -1)
		const int w0 = n > 0 && wv ? wv[0] : w_all;
		KsJob *jobs = pl->jobs;
		auto range = [&](int t, int64_t &lo, int64_t &hi) { lo = n * t / T; hi = n * (t + 1) / T; };
		auto pass1 = [&](int t) { int64_t lo, hi, a = 0, b = 0, c2 = 0; int m = 1, m2 = 1; range(t, lo, hi);
			for (int64_t i = lo; i < hi; ++i) {
				const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
				if (ql != q0len || tl != t0len || (wv && wv[i] != w0)) uni[t] = 0;
				if (ql <= 0 || tl <= 0 || !ok) continue;
				const int tl_ = (tl + 15) / 16;
				a += (int64_t)tl_ * 16; b += (int64_t)ks_qenc_bytes(ql); if (approx) c2 += (int64_t)(extf ? ks_extf2_scratch_bytes(ql, tl) : gg2 ? ks_gg2_scratch_bytes(tl) : ks_scalar_scratch_bytes(tl)); m = std::max(m, tl_); m2 = std::max(m2, ql);
			}
			te[t + 1] = a; qe[t + 1] = b; sc[t + 1] = c2; mt[t] = m; mq[t] = m2; };
		auto pass2 = [&](int t) { int64_t lo, hi; range(t, lo, hi); int64_t a = te[t], b = qe[t], c2 = sc[t];
			for (int64_t i = lo; i < hi; ++i) {
				KsJob &j = jobs[i];
				j.qoff = qoff[i]; j.toff = toff[i]; j.qlen = (int32_t)(qoff[i + 1] - qoff[i]); j.tlen = (int32_t)(toff[i + 1] - toff[i]);
				j.idx = (int32_t)i; j.poff = 0; j.teoff = j.qeoff = j.soff = 0; j.w = ks_eff_w(kind, wv ? wv[i] : w_all, j.qlen, j.tlen);
				if (j.qlen <= 0 || j.tlen <= 0 || !ok) continue;
				j.teoff = a; a += (int64_t)((j.tlen + 15) / 16) * 16;
				j.qeoff = b; b += (int64_t)ks_qenc_bytes(j.qlen);
				if (approx) { j.soff = c2; c2 += (int64_t)(extf ? ks_extf2_scratch_bytes(j.qlen, j.tlen) : gg2 ? ks_gg2_scratch_bytes(j.tlen) : ks_scalar_scratch_bytes(j.tlen)); }
			} };
		auto run = [&](const std::function<void(int)> &f) {
			if (T == 1) { f(0); return; }
			std::vector<std::thread> th; for (int t = 0; t < T; ++t) th.emplace_back(f, t); for (auto &x : th) x.join(); };
		run(pass1);
		for (int t = 0; t < T; ++t) { te[t + 1] += te[t]; qe[t + 1] += qe[t]; sc[t + 1] += sc[t]; pl->max_tlen_ = std::max(pl->max_tlen_, mt[t]); pl->max_qlen = std::max(pl->max_qlen, mq[t]); }
		pl->tenc_bytes = te[T]; pl->qenc_bytes = qe[T]; pl->scal_bytes = sc[T];
		if (pl->rows || pl->extf || pl->gg2) pl->tenc_bytes = pl->qenc_bytes = 0;       // these kernels read the raw sequences
		all_uniform = true; for (int t = 0; t < T; ++t) if (!uni[t]) all_uniform = false;
		if (all_uniform && ok && !pl->cig && n > 0 && q0len > 0 && t0len > 0) {
			pl->uniform = true; pl->u_qlen = q0len; pl->u_tlen = t0len; pl->u_q0 = qoff[0]; pl->u_t0 = toff[0]; pl->u_w = ks_eff_w(kind, w0, q0len, t0len);
			pl->u_te = te[T] / n; pl->u_qe = qe[T] / n; pl->u_sc = sc[T] / n;
		} else run(pass2);
	}
	// 2) per segment: sort (only when lengths differ) so that the 32 jobs of a warp share a geometry; cut into chunks that fit the direction arena
	for (int sg = 0; sg < nseg; ++sg) {
		Seg S; S.lo = bounds[sg]; S.hi = bounds[sg + 1]; S.c0 = pl->chunks.size();
		bool uniform = true;
		for (int64_t i = S.lo + 1; i < S.hi && uniform && !all_uniform; ++i)
			if (pl->jobs[i].qlen != pl->jobs[S.lo].qlen || pl->jobs[i].tlen != pl->jobs[S.lo].tlen) uniform = false;
		if (!uniform)
			std::sort(pl->jobs + S.lo, pl->jobs + S.hi, [](const KsJob &a, const KsJob &b) {
				if (a.tlen != b.tlen) return a.tlen > b.tlen;
				if (a.qlen != b.qlen) return a.qlen > b.qlen;
				return a.idx < b.idx; });
		Chunk cur = {S.lo, S.lo, 0, 0, sg, false, false, false, 1, false};
		if (pl->cig && pl->prep == KS_PREP_OK) {
			// balanced chunks: as few as the arena budget allows, all about the same size (a small last chunk would run at low occupancy)
			auto words_of = [&](const KsJob &j) { const int w = j.w;
				if (pl->rows) return (int64_t)((ks_rows_z_bytes(pl->RP, j.qlen, j.tlen) + 15) / 16);
				if (pl->gg2) return (int64_t)((ks_gg2_dir_bytes(pl->GP, j.qlen, j.tlen) + 15) / 16);
				return (int64_t)((j.tlen + 15) / 16) * ks_prows(j.qlen, j.tlen, w); };
			int64_t total = 0, totc = 0;
			for (int64_t i = S.lo; i < S.hi; ++i) { const KsJob &j = pl->jobs[i]; if (j.qlen > 0 && j.tlen > 0) { total += words_of(j); totc += (int64_t)j.qlen + j.tlen + 1; } }
			const int64_t arena_words_max = std::max<int64_t>(1, arena_words_max_for(total));
			const int64_t nck = std::max<int64_t>(1, std::max((total + arena_words_max - 1) / arena_words_max, (totc + cig_words_max - 1) / cig_words_max));
			const int64_t target = std::min(arena_words_max, total / nck + total / (nck * 64) + 1), ctarget = std::min(cig_words_max, totc / nck + totc / (nck * 64) + 1);
			for (int64_t i = S.lo; i < S.hi; ++i) {
				KsJob &j = pl->jobs[i];
				if (j.qlen <= 0 || j.tlen <= 0) { cur.hi = i + 1; continue; }
				const int64_t words = words_of(j), cc = (int64_t)j.qlen + j.tlen + 1;
				if (cur.hi > cur.lo && (cur.pwords + words > target || cur.cigcap + cc > ctarget)) {
					pl->chunks.push_back(cur); cur.lo = cur.hi = i; cur.pwords = cur.cigcap = 0;
				}
				j.poff = cur.pwords; cur.pwords += words; cur.cigcap += cc;
				cur.hi = i + 1;
			}
		}
		cur.hi = S.hi;
		if (cur.hi > cur.lo) pl->chunks.push_back(cur);
		S.c1 = pl->chunks.size();
		pl->segs.push_back(S);
	}
	// scratch sizing
	// short pairs (the saved state of all resident threads is about the size of L2): do not save the coded target word (KsParams::treload)
	pl->P.treload = (getenv("KSW2B_TRELOAD") ? atoi(getenv("KSW2B_TRELOAD")) != 0 : pl->max_tlen_ <= 32) ? 1 : 0;
	const int SW = ks_save_words(pl->P, pl->P.kind == KS_Z ? (int)KsSaveWords<KS_Z>::value : pl->P.kind == KS_D ? (int)KsSaveWords<KS_D>::value : (int)KsSaveWords<KS_S>::value);
	// One thread per pair needs ~ (SMs x CTAs x threads) concurrent pairs.  A launch (chunk) that cannot fill the GPU that way runs one WARP per
	// pair: few long pairs (>= 24 blocks: a wave of 32 lanes is mostly busy) -- the direction arena bounds the pairs in flight of long CIGAR
	// pairs, and a batch of mixed lengths is sorted by length, so its chunks of long pairs are exactly that case -- or so few pairs that every
	// pair can have a resident warp of its own: then the warp's wavefront cuts the latency of the launch (a lone 150 bp pair: ~1 ms on one
	// thread; this is what the combining layer of the single-pair API sees).  Decided per chunk.
	const int64_t thread_slots = (int64_t)ctx->num_sm * ctx->ctas_per_sm * ctx->threads;
	const bool tiles = !pl->rows && !pl->extf && !pl->gg2 && !pl->approx;
	// Tail split.  A chunk of MIXED lengths that the arena did not have to cut (a rank's share of a strong-scaled batch) would run one thread per
	// pair for as long as its longest pair takes on ONE thread (a 20 kb banded pair: ~1 s) with most of the GPU idle.  Pairs are sorted by length
	// (descending) inside a segment, so the head of such a chunk is split off into a chunk of its own that the rules below put on the ring
	// schedule (one warp per pair: ~9 x faster per pair).  The split point minimises a two-term time model built from this round's measurements:
	// thread mode 16 MCUPS per pair in flight up to the GPU's thread slots and never faster than its longest pair, ring 190 GCUPS (160 MCUPS per pair).
	if (tiles && ctx->mode == 0 && pl->P.kind != KS_S && !pl->uniform && pl->prep == KS_PREP_OK) {
		std::vector<Chunk> out;
		std::vector<double> pre;
		for (const Chunk &c0 : pl->chunks) {
			const int64_t np = c0.hi - c0.lo;
			bool split = false;
			if (np >= 256) {
				pre.assign((size_t)np + 1, 0.0);
				int64_t elig = 0;                                   // pairs of the head that the ring can take (band <= 512, >= 20 blocks of it)
				bool head = true;
				for (int64_t i = 0; i < np; ++i) {
					const KsJob &j = pl->jobs[c0.lo + i];
					pre[(size_t)i + 1] = pre[(size_t)i] + (double)pair_cost_w(j.w, j.qlen, j.tlen, pl->cig != 0);
					if (head && j.w <= KS_RING_MAX_W && std::min((j.tlen + 15) / 16, (j.w + 16) / 16 + 1) >= 20) elig = i + 1; else head = false;
				}
				const double r1 = 16e6, ring_max = 190e9 * (pl->P.kind == KS_Z ? 1.5 : 1.0), ring_1 = ring_max / 1184.0;
				auto t_thread = [&](int64_t from) -> double { const int64_t m = np - from; if (m <= 0) return 0.0;
					const double S = pre[(size_t)np] - pre[(size_t)from], cmax = pre[(size_t)from + 1] - pre[(size_t)from];
					return std::max(S / ((double)std::min<int64_t>(m, thread_slots) * r1), cmax / r1); };
				auto t_ring = [&](int64_t i) -> double { return i <= 0 ? 0.0 : pre[(size_t)i] / std::min(ring_max, (double)i * ring_1); };
				const double t0 = t_thread(0);
				double best_t = t0; int64_t best_i = 0;
				for (int64_t i = 64; i <= elig; i = i + std::max<int64_t>(64, i / 8)) { const double t = t_ring(i) + t_thread(i); if (t < best_t) { best_t = t; best_i = i; } }
				if (best_i > 0 && best_i < np && best_t < 0.8 * t0) {
					Chunk a = c0, b = c0;
					a.hi = c0.lo + best_i; b.lo = a.hi;
					a.cigcap = b.cigcap = 0;
					for (int64_t i = a.lo; i < a.hi; ++i) if (pl->jobs[i].qlen > 0 && pl->jobs[i].tlen > 0) a.cigcap += (int64_t)pl->jobs[i].qlen + pl->jobs[i].tlen + 1;
					for (int64_t i = b.lo; i < b.hi; ++i) if (pl->jobs[i].qlen > 0 && pl->jobs[i].tlen > 0) b.cigcap += (int64_t)pl->jobs[i].qlen + pl->jobs[i].tlen + 1;
					a.force_ring = true;
					out.push_back(a); out.push_back(b);                 // (both keep their offsets into the same direction arena: the original chunk fits it)
					split = true;
				}
			}
			if (!split) out.push_back(c0);
		}
		if (out.size() != pl->chunks.size()) {
			pl->chunks.swap(out);
			for (Seg &S : pl->segs) { S.c0 = pl->chunks.size(); S.c1 = 0; }
			for (size_t ci = 0; ci < pl->chunks.size(); ++ci) { Seg &S = pl->segs[(size_t)pl->chunks[ci].seg]; S.c0 = std::min(S.c0, ci); S.c1 = std::max(S.c1, ci + 1); }
		}
	}
	int64_t big_thread = 0, big_warp = 0, big_cta = 0;
	int mt_thread = 1, mt_warp = 1;
	pl->warp_mode = false;
	for (auto &c : pl->chunks) {
		const int64_t np = c.hi - c.lo;
		c.max_tlen_ = 1;
		if (pl->uniform) c.max_tlen_ = pl->max_tlen_;
		else for (int64_t i = c.lo; i < c.hi; ++i) c.max_tlen_ = std::max(c.max_tlen_, (pl->jobs[i].tlen + 15) / 16);
		// blocks a diagonal of the band spans (what a wave of lanes can be busy with): the widest of the chunk
		int band_blocks = 0, max_w = 0;
		if (pl->uniform) { band_blocks = std::min(pl->max_tlen_, (pl->u_w + 16) / 16 + 1); max_w = pl->u_w; }
		else for (int64_t i = c.lo; i < c.hi; ++i) { band_blocks = std::max(band_blocks, std::min((pl->jobs[i].tlen + 15) / 16, (pl->jobs[i].w + 16) / 16 + 1)); max_w = std::max(max_w, pl->jobs[i].w); }
		// Few pairs with a WIDE band (>= 48 blocks: the 32 lanes of a wave stay busy; measured on 33-block bands the wavefront is only ~40 % occupied and
		// the warp kernel runs at 0.4 x the thread kernel even when that one is short of pairs), or so few pairs that each can have a warp of its own
		c.warp = tiles && (ctx->mode == 2 || ctx->mode == 3 || (ctx->mode == 0 && ((np * 3 < thread_slots && band_blocks >= 48) || (np * 32 <= thread_slots && c.max_tlen_ >= 3))));
		// Banded pairs (effective band <= 512, at least ~20 blocks of it) that are FAR from filling the GPU with a thread each -- the longest CIGAR pairs,
		// whose direction bytes (20 MB for a 20 kb pair at w = 500) bound the pairs in flight: one warp per pair on the ring schedule.  Measured on
		// 5 kb dual-gap CIGAR pairs: ring 193 GCUPS whatever the number of pairs, thread mode 18 MCUPS per pair in flight (362 GCUPS at 20 k pairs):
		// the ring wins below ~11 k pairs per launch.  (exts2 has no band: never.)
		c.ring = tiles && pl->P.kind != KS_S && max_w <= KS_RING_MAX_W && np > 0 &&
		         (ctx->mode == 4 || c.force_ring || (ctx->mode == 0 && !c.warp && band_blocks >= 20 && np * 5 < thread_slots));
		if (c.ring) c.warp = true;
		// very few pairs whose band is hundreds of blocks wide: one CTA per pair (exact-max kernels only)
		c.cta = c.warp && !c.ring && !(pl->P.flag & KSF_APPROX_MAX) && np > 0 && (ctx->mode == 3 || (ctx->mode == 0 && np * 2 <= ctx->num_sm && band_blocks >= KS_CTA_LANES / 2));
		if (c.cta) { big_cta = std::max(big_cta, np); mt_warp = std::max(mt_warp, c.max_tlen_); pl->warp_mode = true; }
		else if (c.warp) { big_warp = std::max(big_warp, np); mt_warp = std::max(mt_warp, c.max_tlen_); pl->warp_mode = true; }
		else { big_thread = std::max(big_thread, np); mt_thread = std::max(mt_thread, c.max_tlen_); }
	}
	const int warps_per_cta = ctx->threads / 32;
	pl->grid_warp = (int)std::max<int64_t>(1, std::min<int64_t>(std::max((big_warp + 3) / 4, big_cta), (int64_t)ctx->num_sm * 4));   // (CTA mode: one pair per CTA, 4 x fewer stream / save slots used)
	pl->grid = (int)std::max<int64_t>(1, std::min<int64_t>((big_thread + 32ll * warps_per_cta - 1) / (32ll * warps_per_cta), (int64_t)ctx->num_sm * ctx->ctas_per_sm));
	// save area: one slot of max_tlen_ blocks per resident thread (thread mode) or warp (warp mode); both kinds of chunk share it
	size_t save_need = 0;
	if (big_thread > 0) save_need = std::max(save_need, ((size_t)pl->grid * ctx->threads + 32) * (size_t)mt_thread * SW);
	if (big_warp > 0 || big_cta > 0) save_need = std::max(save_need, ((size_t)pl->grid_warp * 4 + 32) * (size_t)mt_warp * SW);
	pl->save_stride_thread = (size_t)mt_thread * SW; pl->save_stride_warp = (size_t)mt_warp * SW;
	if (pl->extf || pl->gg2) { pl->warp_mode = false; save_need = 0; }
	pl->wpanel = std::max(1, std::min(ctx->wpanel, pl->max_qlen + 16 * pl->max_tlen_));     // (no panel is taller than the longest pair's diagonals)
	if (pl->rows) {                                        // one scratch slot per resident warp, sized for the longest query; at most ~4 GiB in all
		pl->warp_mode = false; save_need = 0;
		pl->rows_warp_words = 32 * ks_rows_eh_words(pl->max_qlen);
		const int64_t fit = std::max<int64_t>(1, (int64_t)((4ull << 30) / (pl->rows_warp_words * 4 * 4)));
		pl->rows_grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((n + 127) / 128, (int64_t)ctx->num_sm * 4), fit));
		pl->scal_bytes = (int64_t)pl->rows_grid * 4 * (int64_t)pl->rows_warp_words * 4;
	}
	int64_t max_p = 0, max_c = 0;
	for (auto &c : pl->chunks) { max_p = std::max(max_p, c.pwords); max_c = std::max(max_c, c.cigcap); }
	if (ctx->d_jobs.ensure(sizeof(KsJob) * (size_t)std::max<int64_t>(1, n)) || ctx->d_res.ensure(sizeof(KsResult) * (size_t)std::max<int64_t>(1, n)) ||
	    ctx->d_save.ensure((pl->cig ? 1 : 2) * (pl->save_words = save_need) * 16) || ctx->d_ctr.ensure(4096) ||
	    (pl->warp_mode && ctx->d_wv.ensure((pl->cig ? 1 : 2) * (pl->wv_words = ((size_t)pl->grid_warp * 4 + 4) * KS_WARP_WV_WORDS(pl->wpanel)) * 16)) ||
	    ctx->d_tenc.ensure((size_t)pl->tenc_bytes + 64) || ctx->d_qenc.ensure((size_t)pl->qenc_bytes + 64) || ctx->d_scal.ensure((size_t)pl->scal_bytes + 64) ||
	    (pl->cig && (ctx->d_parena.ensure((size_t)std::max<int64_t>(1, max_p) * 16) || ctx->d_cig.ensure((size_t)std::max<int64_t>(1, max_c) * 4 * (pl->chunks.size() > 1 ? 2 : 1))))) {
		ks_fail(-11, "device allocation failed (jobs %lld, save %zu B, arena %lld B)", (long long)n, save_need * 16, (long long)max_p * 16);
		delete pl; return 0;
	}
	pl->cig_half = pl->chunks.size() > 1 ? std::max<int64_t>(1, max_c) : 0;
	if (upload && n > 0) {
		if (pl->uniform) {
			if (upload_jobs(pl, 0, n, 0) || cudaStreamSynchronize(0) != cudaSuccess) { ks_fail(-10, "job table generation failed"); delete pl; return 0; }
		} else if (cudaMemcpy(ctx->d_jobs.p, pl->jobs, sizeof(KsJob) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
			ks_fail(-10, "job table upload failed"); delete pl; return 0;
		}
	}
	return pl;
}

extern "C" ksw2b_plan_t *ksw2b_plan_create_ex(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const int64_t *qoff, const int64_t *toff, const int32_t *w)
{
	if (n < 0) { ks_fail(-2, "bad arguments"); return 0; }
	return plan_build(ctx, par, n, qoff, toff, w, std::vector<int64_t>{0, n}, true);
}
extern "C" ksw2b_plan_t *ksw2b_plan_create(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const int64_t *qoff, const int64_t *toff)
{
	return ksw2b_plan_create_ex(ctx, par, n, qoff, toff, 0);
}

// Opt a kernel in to `smem` bytes of dynamic shared memory -- not on every launch, and only ever UPWARDS: the attribute belongs to the
// (device, kernel) pair, not to a context, so the record of what has been granted is process-wide (a second context lowering it would
// make the first context's launches fail with "invalid argument").
static int ks_optin_smem(ksw2b_ctx *ctx, const void *fn, size_t smem)
{
	static std::mutex mu;
	static std::unordered_map<uint64_t, size_t> granted;
	std::lock_guard<std::mutex> lk(mu);
	const uint64_t key = (uint64_t)(uintptr_t)fn * 64u + (uint64_t)(ctx->device & 63);
	auto it = granted.find(key);
	if (it != granted.end() && it->second >= smem) return 0;
	CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	granted[key] = smem;
	return 0;
}

// (exts2 has no band: no ring kernels are instantiated for it)
template<int KIND, int CIG>
static void ks_launch_ring(ksw2b_plan *pl, const Chunk &ch, int grid, long long nj, const uint8_t *dq, const uint8_t *dt, const uint8_t *dj, unsigned long long *ctr, cudaStream_t st)
{
	ksw2b_ctx *ctx = pl->ctx;
	if constexpr (KIND != KS_S)
		ks_fill_ring_kernel<KIND, CIG><<<grid, 128, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, ctr, dq, dt, dj, (const uint8_t*)ctx->d_tenc.p, (const uint8_t*)ctx->d_qenc.p,
		                                                   (ks_u4*)ctx->d_save.p + (size_t)pl->slot * pl->save_words, pl->save_stride_warp, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p);
}

template<int KIND, int CIG>
static int launch_fill(ksw2b_plan *pl, const Chunk &ch, const uint8_t *dq, const uint8_t *dt, const uint8_t *dj, unsigned long long *ctr, cudaStream_t st)
{
	ksw2b_ctx *ctx = pl->ctx;
	if (ch.cta && KS_APX(CIG) == 0) {
		const int C = pl->wpanel;
		const long long nj = ch.hi - ch.lo;
		const int grid = (int)std::max<long long>(1, std::min<long long>(nj, pl->grid_warp));
		ks_fill_cta_kernel<KIND, KS_DIR(CIG)><<<grid, KS_CTA_LANES, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, ctr, dq, dt, dj,
		                                                      (const uint8_t*)ctx->d_tenc.p, (const uint8_t*)ctx->d_qenc.p,
		                                                      (ks_u4*)ctx->d_save.p + (size_t)pl->slot * pl->save_words, pl->save_stride_warp,
		                                                      (ks_u4*)ctx->d_wv.p + (size_t)pl->slot * pl->wv_words, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p, C);
		CK(cudaGetLastError());
		return 0;
	}
	if (ch.ring) {
		const long long nj = ch.hi - ch.lo;
		const int grid = (int)std::max<long long>(1, std::min<long long>((nj + 3) / 4, pl->grid_warp));
		ks_launch_ring<KIND, CIG>(pl, ch, grid, nj, dq, dt, dj, ctr, st);
		CK(cudaGetLastError());
		return 0;
	}
	if (ch.warp) {
		const int C = pl->wpanel;
		const long long nj = ch.hi - ch.lo;
		const int grid = (int)std::max<long long>(1, std::min<long long>((nj + 3) / 4, pl->grid_warp));
		ks_fill_warp_kernel<KIND, CIG><<<grid, 128, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, ctr, dq, dt, dj,
		                                                      (const uint8_t*)ctx->d_tenc.p, (const uint8_t*)ctx->d_qenc.p,
		                                                      (ks_u4*)ctx->d_save.p + (size_t)pl->slot * pl->save_words, pl->save_stride_warp,
		                                                      (ks_u4*)ctx->d_wv.p + (size_t)pl->slot * pl->wv_words, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p, C);
		CK(cudaGetLastError());
		return 0;
	}
	const long long nj = ch.hi - ch.lo;
	// Launch shape.  A launch that fills the GPU uses the tuned CTA size; one that cannot (few long pairs: the direction arena bounds
	// the pairs in flight) runs ONE WARP PER CTA so that the block scheduler spreads the warps evenly over the SMs (209 CTAs of 96
	// threads on 148 SMs leave 87 SMs with half the work of the other 61).
	int tpb = std::min(ctx->threads, KS_MAX_TPB(KIND));
	const long long warps_needed = (nj + 31) / 32, warp_slots = (long long)pl->grid * (ctx->threads / 32);
	long long grid_ll = std::min<long long>((warps_needed + tpb / 32 - 1) / (tpb / 32), pl->grid);
	if (warps_needed <= warp_slots && ctx->auto_panel) { tpb = 32; grid_ll = warps_needed; }
	const int grid = (int)std::max<long long>(1, grid_ll);
	// Panel height: the tuned default fills the SM's shared memory at full occupancy.  An under-filled launch, or a kernel whose
	// registers let fewer CTAs be resident (the dual-gap / splice kernels need ~190: three 96-thread CTAs per SM), gets taller panels
	// from the shared memory that is left -- fewer tile save / restore round trips through L2 (profiles/r1_tuning.txt).
	int C = ctx->panel;
	cudaFuncAttributes fa;
	{ auto it = ctx->fattr.find((const void*)ks_fill_kernel<KIND, CIG>);
	  if (it == ctx->fattr.end()) { CK(cudaFuncGetAttributes(&fa, ks_fill_kernel<KIND, CIG>)); ctx->fattr[(const void*)ks_fill_kernel<KIND, CIG>] = fa; } else fa = it->second; }
	const int by_regs = std::max(1, 65536 / (std::max(1, fa.numRegs) * tpb));
	const int resident = std::max(1, std::min((grid + ctx->num_sm - 1) / ctx->num_sm, by_regs));
	if (ctx->auto_panel && resident * tpb < ctx->ctas_per_sm * ctx->threads) {
		const long long budget = (long long)ctx->smem_sm / resident - 1024;
		const int tall = (int)std::min<long long>(36, (budget / (16ll * tpb) - 1) / 2);
		C = std::max(C, tall);
	}
	// The approximate-max kernels keep no arg-max stream (one record per diagonal instead of two): twice the panel height in the same shared memory,
	// half the tile save / restore round trips -- that mode runs so few instructions per step that the L2 / DRAM latency of the restores shows
	// (long-scoreboard stalls 1.26 per issue at panel 15).
	const int per_diag = KS_APX(CIG) ? 1 : 2;
	if (KS_APX(CIG)) C = 2 * C;
	size_t smem = (size_t)(per_diag * C + 1) * 16 * tpb;
	if (smem > ctx->smem_optin && C > ctx->panel * (3 - per_diag)) { C = ctx->panel * (3 - per_diag); smem = (size_t)(per_diag * C + 1) * 16 * tpb; }     // the tall panel does not fit this device: the tuned default
	if (smem > ctx->smem_optin) return ks_fail(-12, "panel %d x %d threads needs %zu B shared memory (max %zu)", C, tpb, smem, ctx->smem_optin);
	{ int rc = ks_optin_smem(ctx, (const void*)ks_fill_kernel<KIND, CIG>, smem); if (rc) return rc; }
	if (ctx->l2_persist && ctx->l2_window_max > 0 && ctx->l2_persist_max > 0) {
		// The saved block state of all resident threads (~100 MB on the 150 bp workload) is about the size of L2 and is rewritten every panel, while
		// 600 MB of sequences stream through the same cache per launch: mark the save area persisting so that the stream does not evict it.
		const char *base = (const char*)((ks_u4*)ctx->d_save.p + (size_t)pl->slot * pl->save_words);
		const size_t used = std::min<size_t>((size_t)grid * tpb * pl->save_stride_thread * 16, ctx->l2_window_max);
		const int si = pl->slot & 1;
		if (ctx->l2_win_ptr[si] != base || ctx->l2_win_bytes[si] != used || ctx->l2_win_stream[si] != st) {
			cudaStreamAttrValue av; memset(&av, 0, sizeof av);
			av.accessPolicyWindow.base_ptr = (void*)base; av.accessPolicyWindow.num_bytes = used;
			av.accessPolicyWindow.hitRatio = used > ctx->l2_persist_max ? (float)((double)ctx->l2_persist_max / (double)used) : 1.0f;
			av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
			ctx->l2_win_ptr[si] = base; ctx->l2_win_bytes[si] = used; ctx->l2_win_stream[si] = st;
		}
	}
	ks_fill_kernel<KIND, CIG><<<grid, tpb, smem, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, ctr, dq, dt, dj,
	                                                   (const uint8_t*)ctx->d_tenc.p, (const uint8_t*)ctx->d_qenc.p,
	                                                   (ks_u4*)ctx->d_save.p + (size_t)pl->slot * pl->save_words, pl->save_stride_thread, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p, C);
	CK(cudaGetLastError());
	return 0;
}

static int launch_fill_any(ksw2b_plan *pl, const Chunk &ch, const uint8_t *dq, const uint8_t *dt, const uint8_t *dj, unsigned long long *ctr, cudaStream_t st)
{
#define GO(K, G) return launch_fill<K, G>(pl, ch, dq, dt, dj, ctr, st)
	const int k = pl->P.kind, g = pl->cig;
	if (pl->P.flag & KSF_APPROX_MAX) {             // the approximate-max kernel variants (template argument CIG + 4)
		if (k == KS_Z) { if (g == 0) GO(KS_Z, 4); if (g == 1) GO(KS_Z, 5); GO(KS_Z, 6); }
		if (k == KS_D) { if (g == 0) GO(KS_D, 4); if (g == 1) GO(KS_D, 5); GO(KS_D, 6); }
		if (g == 0) GO(KS_S, 4); if (g == 1) GO(KS_S, 5); GO(KS_S, 6);
	}
	if (k == KS_Z) { if (g == 0) GO(KS_Z, 0); if (g == 1) GO(KS_Z, 1); GO(KS_Z, 2); }
	if (k == KS_D) { if (g == 0) GO(KS_D, 0); if (g == 1) GO(KS_D, 1); GO(KS_D, 2); }
	if (g == 0) GO(KS_S, 0); if (g == 1) GO(KS_S, 1); GO(KS_S, 2);
#undef GO
}

// results of pairs that never reach a kernel (invalid parameters): what the reference leaves after ksw_reset_extz
static void fill_reset(ksw2b_result_t *r)
{
	memset(r, 0, sizeof *r);
	r->max_q = r->max_t = r->mqe_t = r->mte_q = -1; r->mqe = r->mte = r->score = KS_NEG_INF; r->tb_i = r->tb_j = -1;
}

// host copy of chunk ci's CIGAR words (several-chunk CIGAR runs): waits for the chunk's traceback, appends the words to ctx->cig_host
static int drain_chunk_cigars(ksw2b_plan *pl, size_t ci)
{
	ksw2b_ctx *ctx = pl->ctx;
	unsigned long long used = 0;
	CK(cudaEventSynchronize(pl->cig_ev[ci & 1]));
	CK(cudaMemcpy(&used, (unsigned long long*)ctx->d_ctr.p + 2 * (ci % 64) + 1, 8, cudaMemcpyDeviceToHost));
	const size_t old = ctx->cig_host.size();
	ctx->cig_host.resize(old + (size_t)used);
	if (used) CK(cudaMemcpy(ctx->cig_host.data() + old, (uint32_t*)ctx->d_cig.p + (int64_t)(ci & 1) * pl->cig_half, (size_t)used * 4, cudaMemcpyDeviceToHost));
	pl->chunk_cig_used[ci] = (int64_t)used;
	pl->drained = (int64_t)ci + 1;
	return 0;
}

// One chunk on stream st: encode -> fill (or scalar) -> traceback.  ci: chunk index.  In CIGAR mode with several chunks the
// chunk's CIGAR words are drained to the host before the next chunk re-uses the staging buffer and the direction arena.
static int run_chunk(ksw2b_plan *pl, size_t ci, const uint8_t *d_qcat, const uint8_t *d_tcat, const uint8_t *d_junc, cudaStream_t st)
{
	ksw2b_ctx *ctx = pl->ctx;
	const Chunk &ch = pl->chunks[ci];
	if (ch.hi <= ch.lo) return 0;
	const long long nj = ch.hi - ch.lo;
	unsigned long long *ctrs = (unsigned long long*)ctx->d_ctr.p + 2 * (ci % 64);
	CK(cudaMemsetAsync(ctrs, 0, 16, st));
	if (pl->gg2) {
		ks_gg2_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->GP, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, d_qcat, d_tcat, (int8_t*)ctx->d_scal.p,
		                                                       (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p, pl->cig ? 1 : 0);
		CK(cudaGetLastError());
		++pl->launches;
	} else if (pl->extf) {
		ks_extf2_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->FP, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, d_qcat, d_tcat, (uint8_t*)ctx->d_scal.p, (KsResult*)ctx->d_res.p);
		CK(cudaGetLastError());
		++pl->launches;
	} else if (pl->rows) {
		ks_rows_kernel<<<pl->rows_grid, 128, 0, st>>>(pl->RP, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, ctrs, d_qcat, d_tcat, (int32_t*)ctx->d_scal.p,
		                                               pl->rows_warp_words, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p);
		CK(cudaGetLastError());
		++pl->launches;
	} else if (pl->approx) {
		ks_scalar_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, d_qcat, d_tcat, d_junc,
		                                                          (int8_t*)ctx->d_scal.p, (ks_u4*)ctx->d_parena.p, (KsResult*)ctx->d_res.p);
		CK(cudaGetLastError());
		++pl->launches;
	} else {
		ks_encode_kernel<<<(unsigned)((nj * 32 + 255) / 256), 256, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, d_qcat, d_tcat, (uint8_t*)ctx->d_tenc.p, (uint8_t*)ctx->d_qenc.p);
		CK(cudaGetLastError());
		if (pl->timing) {
			while (pl->tev.size() < pl->tev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); pl->tev.push_back(e); }
			CK(cudaEventRecord(pl->tev[pl->tev_used], st));
		}
		int rc = launch_fill_any(pl, ch, d_qcat, d_tcat, d_junc, ctrs, st);
		if (rc) return rc;
		if (pl->timing) { CK(cudaEventRecord(pl->tev[pl->tev_used + 1], st)); pl->tev_used += 2; }
		pl->launches += 2;
	}
	if (pl->cig) {
		if (pl->gg2)
			ks_gg2_traceback_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->GP, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, (const ks_u4*)ctx->d_parena.p,
			                                                                 (KsResult*)ctx->d_res.p, (uint32_t*)ctx->d_cig.p + (int64_t)(ci & 1) * pl->cig_half, ctrs + 1, ch.cigcap);
		else if (pl->rows)
			ks_rows_traceback_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->RP, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, (const ks_u4*)ctx->d_parena.p,
			                                                                  (KsResult*)ctx->d_res.p, (uint32_t*)ctx->d_cig.p + (int64_t)(ci & 1) * pl->cig_half, ctrs + 1, ch.cigcap);
		else
		ks_traceback_kernel<<<(unsigned)((nj + 63) / 64), 64, 0, st>>>(pl->P, (const KsJob*)ctx->d_jobs.p + ch.lo, nj, d_qcat, d_tcat, (const ks_u4*)ctx->d_parena.p,
		                                                             (KsResult*)ctx->d_res.p, (uint32_t*)ctx->d_cig.p + (int64_t)(ci & 1) * pl->cig_half, ctrs + 1, ch.cigcap);
		CK(cudaGetLastError());
		++pl->launches;
		if (pl->chunks.size() > 1) {
			// The chunk's CIGAR words are fetched with ONE CHUNK OF LAG: this chunk's kernels are queued first, then the host drains the previous
			// chunk's half of the staging buffer while the GPU fills this one (the direction arena is shared, so fill(k+1) follows traceback(k)
			// on the stream; only the host copy used to sit between them).
			if (!pl->cig_ev[ci & 1]) CK(cudaEventCreateWithFlags(&pl->cig_ev[ci & 1], cudaEventDisableTiming));
			CK(cudaEventRecord(pl->cig_ev[ci & 1], st));
			while (pl->drained < (int64_t)ci) { int rc = drain_chunk_cigars(pl, (size_t)pl->drained); if (rc) return rc; }
		}
	}
	return 0;
}

extern "C" int ksw2b_plan_run(ksw2b_plan_t *pl, const uint8_t *d_qcat, const uint8_t *d_tcat, const uint8_t *d_junc, void *stream)
{
	if (!pl) return ks_fail(-2, "null plan");
	ksw2b_ctx *ctx = pl->ctx;
	cudaStream_t st = (cudaStream_t)stream;
	CK(cudaSetDevice(ctx->device));
	pl->launches = 0; pl->ran = true; pl->tev_used = 0;
	pl->chunk_cig_used.assign(pl->chunks.size(), 0);
	if (pl->prep != KS_PREP_OK || pl->n == 0) return 0;
	if (pl->chunks.size() > 1 && pl->cig) ctx->cig_host.clear();
	pl->drained = 0;
	for (size_t ci = 0; ci < pl->chunks.size(); ++ci) { int rc = run_chunk(pl, ci, d_qcat, d_tcat, d_junc, st); if (rc) return rc; }
	return 0;
}

// CIGAR words + per-pair offsets after all chunks ran on stream st (results already on the host in res[])
static int collect_cigars(ksw2b_plan *pl, ksw2b_result_t *res, const uint32_t **cigar, cudaStream_t st)
{
	ksw2b_ctx *ctx = pl->ctx;
	if (pl->chunks.size() == 1) {
		unsigned long long used = 0;
		CK(cudaMemcpyAsync(&used, (unsigned long long*)ctx->d_ctr.p + 1, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		ctx->cig_host.resize((size_t)used);
		if (used) CK(cudaMemcpyAsync(ctx->cig_host.data(), ctx->d_cig.p, (size_t)used * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	} else {
		while (pl->drained < (int64_t)pl->chunks.size()) { int rc = drain_chunk_cigars(pl, (size_t)pl->drained); if (rc) return rc; }   // (the last chunk; all of them if no chunk ran after)
		CK(cudaStreamSynchronize(st));
		int64_t base = 0;                                   // per-chunk offsets were relative to the chunk's staging buffer: rebase
		for (size_t ci = 0; ci < pl->chunks.size(); ++ci) {
			for (int64_t k = pl->chunks[ci].lo; k < pl->chunks[ci].hi; ++k) res[pl->jobs[k].idx].cigar_off += base;
			base += pl->chunk_cig_used[ci];
		}
	}
	if (cigar) *cigar = ctx->cig_host.data();
	return 0;
}

extern "C" int ksw2b_plan_fetch(ksw2b_plan_t *pl, ksw2b_result_t *res, const uint32_t **cigar, void *stream)
{
	if (!pl || !res) return ks_fail(-2, "bad arguments");
	ksw2b_ctx *ctx = pl->ctx;
	cudaStream_t st = (cudaStream_t)stream;
	CK(cudaSetDevice(ctx->device));
	if (cigar) *cigar = 0;
	if (pl->prep != KS_PREP_OK || pl->n == 0) { for (int64_t i = 0; i < pl->n; ++i) fill_reset(&res[i]); return 0; }
	static_assert(sizeof(KsResult) == sizeof(ksw2b_result_t), "result layouts must match");
	CK(cudaMemcpyAsync(res, ctx->d_res.p, sizeof(KsResult) * (size_t)pl->n, cudaMemcpyDeviceToHost, st));
	if (pl->cig) return collect_cigars(pl, res, cigar, st);
	CK(cudaStreamSynchronize(st));
	return 0;
}

extern "C" const ksw2b_result_t *ksw2b_plan_device_results(ksw2b_plan_t *pl) { return pl ? (const ksw2b_result_t*)pl->ctx->d_res.p : 0; }
extern "C" int64_t ksw2b_plan_cells(ksw2b_plan_t *pl)
{
	if (!pl) return 0;
	if (pl->cells < 0) {                                   // lazily: an O(diagonals) sum per distinct (qlen, tlen)
		struct K3 { int q, t, w; bool operator==(const K3 &o) const { return q == o.q && t == o.t && w == o.w; } };
		struct H3 { size_t operator()(const K3 &k) const { return (((uint64_t)(uint32_t)k.q << 32) | (uint32_t)k.t) * 0x9E3779B97F4A7C15ull + (uint32_t)k.w; } };
		std::unordered_map<K3, int64_t, H3> memo;
		pl->cells = 0;
		if (pl->uniform) {
			const int w = pl->u_w;
			pl->cells = pl->n * (pl->rows ? rows_cells(pl->u_qlen, pl->u_tlen, w) : band_cells(pl->u_qlen, pl->u_tlen, w));
			return pl->cells;
		}
		for (int64_t i = 0; i < pl->n && pl->prep == KS_PREP_OK; ++i) {
			const KsJob &j = pl->jobs[i];
			if (j.qlen <= 0 || j.tlen <= 0) continue;
			const int w = j.w;
			const K3 key = {j.qlen, j.tlen, w};
			auto it = memo.find(key);
			if (it == memo.end()) it = memo.emplace(key, pl->rows ? rows_cells(j.qlen, j.tlen, w) : band_cells(j.qlen, j.tlen, w)).first;
			pl->cells += it->second;
		}
	}
	return pl->cells;
}
extern "C" int ksw2b_plan_launches(ksw2b_plan_t *pl) { return pl ? pl->launches : 0; }
extern "C" void ksw2b_plan_set_timing(ksw2b_plan_t *pl, int on) { if (pl) pl->timing = on != 0; }
// device time of the DP-fill launches of the last ksw2b_plan_run (CUDA events on the launching stream); waits for them
extern "C" double ksw2b_plan_fill_ms(ksw2b_plan_t *pl, int *n_launches)
{
	double ms = 0;
	if (n_launches) *n_launches = 0;
	if (!pl) return 0;
	for (size_t i = 0; i + 1 < pl->tev_used; i += 2) {
		float t = 0;
		if (cudaEventSynchronize(pl->tev[i + 1]) != cudaSuccess || cudaEventElapsedTime(&t, pl->tev[i], pl->tev[i + 1]) != cudaSuccess) { cudaGetLastError(); return -1; }
		ms += t;
		if (n_launches) ++*n_launches;
	}
	return ms;
}
extern "C" void ksw2b_plan_destroy(ksw2b_plan_t *pl) { delete pl; }

// The drop-in batch call: the batch is cut into contiguous segments; segment s+1's sequences and job table travel to the
// device (input stream) while segment s computes (compute streams) and segment s-1's results travel back (output stream).
extern "C" int ksw2b_align_ex(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const uint8_t *qcat, const int64_t *qoff,
                              const uint8_t *tcat, const int64_t *toff, const uint8_t *junc, const int32_t *wv, ksw2b_result_t *res, const uint32_t **cigar)
{
	if (!ctx || !par || !res || n < 0) return ks_fail(-2, "bad arguments");
	CK(cudaSetDevice(ctx->device));
	if (cigar) *cigar = 0;
	if (n == 0) return 0;
	const bool timing = getenv("KSW2B_TIMING") != 0;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
	const double t_start = now();
	// a small first segment gets the GPU busy early; the rest stay large so that kernel tails stay rare
	std::vector<int64_t> bounds{0};
	const int64_t slots = (int64_t)ctx->num_sm * ctx->ctas_per_sm * ctx->threads;      // pairs one launch needs to fill the GPU
	static const int env_two = getenv("KSW2B_TWO_STREAMS") ? atoi(getenv("KSW2B_TWO_STREAMS")) : 1;      // knobs for experiments (profiles/r1_tuning.txt)
	// Segmentation of the pipeline: 10 % / 3 x 28 % / 6 % from pinned caller buffers; from pageable ones (whose copies are staged by the driver and
	// do not overlap as well) more and smaller segments: 5 % / 6 x 15 % / 3 % (measured on the 150 bp workload: 521 -> 568 GCUPS end to end from
	// pageable buffers, and 656 -> 601 from pinned ones: profiles/r2_ab_e2e_segments.txt).  KSW2B_FIRST_PCT / _REST_SEGS / _LAST_PCT override.
	bool in_pinned = true;
	{ cudaPointerAttributes qa; in_pinned = (cudaPointerGetAttributes(&qa, qcat) == cudaSuccess && qa.type == cudaMemoryTypeHost); cudaGetLastError(); }
	const int env_first = getenv("KSW2B_FIRST_PCT") ? atoi(getenv("KSW2B_FIRST_PCT")) : (in_pinned ? 10 : 5), env_rest = getenv("KSW2B_REST_SEGS") ? atoi(getenv("KSW2B_REST_SEGS")) : (in_pinned ? 3 : 6);
	const int env_last = getenv("KSW2B_LAST_PCT") ? atoi(getenv("KSW2B_LAST_PCT")) : (in_pinned ? 6 : 3);
	if (n >= 4 * slots) {
		// small first segment: the GPU starts early; (optional) small last segment: little left to copy back after the last kernel
		const int64_t first = std::max<int64_t>(1, n * std::max(1, std::min(50, env_first)) / 100), last = n * std::max(0, std::min(30, env_last)) / 100, rest = n - first - last;
		const int nrest = std::max(1, std::min(8, env_rest));
		for (int i = 0; i < nrest; ++i) bounds.push_back(first + rest * i / nrest);
		if (last > 0) bounds.push_back(n - last);
	}
	bounds.push_back(n);
	// The first segment's sequences start travelling BEFORE the job table exists: the host builds the plan (threaded, 0.6 - 1.6 ms per
	// million pairs) while they fly.
	const size_t qb = (size_t)qoff[n], tb = (size_t)toff[n];
	if (ctx->d_q.ensure(qb + 64) || ctx->d_t.ensure(tb + 64) || (junc && ctx->d_j.ensure(tb + 64))) return ks_fail(-11, "device allocation failed");
	if (!ctx->s_in) {
		CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ctx->s_job, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&ctx->s_cmp, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ctx->s_cmp2, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
	}
	const size_t nsegs = bounds.size() - 1;
	while (ctx->ev.size() < 4 * nsegs) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev.push_back(e); }
	auto drain = [&]() { cudaStreamSynchronize(ctx->s_in); cudaStreamSynchronize(ctx->s_job); cudaStreamSynchronize(ctx->s_cmp); cudaStreamSynchronize(ctx->s_cmp2); cudaStreamSynchronize(ctx->s_out); };
	// input stream order: seq(0) | host builds the plan meanwhile | jobs(0), seq(1), jobs(1), seq(2), ...  (one stream, so that a segment's job
	// table never queues behind the sequences of LATER segments)
	auto upload_seqs = [&](size_t s) -> cudaError_t {
		const size_t q0 = (size_t)qoff[bounds[s]], q1 = (size_t)qoff[bounds[s + 1]], t0 = (size_t)toff[bounds[s]], t1 = (size_t)toff[bounds[s + 1]];
		cudaError_t e = cudaSuccess;
		if ((q1 > q0 && (e = cudaMemcpyAsync((uint8_t*)ctx->d_q.p + q0, qcat + q0, q1 - q0, cudaMemcpyHostToDevice, ctx->s_in)) != cudaSuccess) ||
		    (t1 > t0 && (e = cudaMemcpyAsync((uint8_t*)ctx->d_t.p + t0, tcat + t0, t1 - t0, cudaMemcpyHostToDevice, ctx->s_in)) != cudaSuccess) ||
		    (junc && t1 > t0 && (e = cudaMemcpyAsync((uint8_t*)ctx->d_j.p + t0, junc + t0, t1 - t0, cudaMemcpyHostToDevice, ctx->s_in)) != cudaSuccess)) return e;
		return cudaSuccess;
	};
	ctx->last_h2d = (unsigned long long)qb + tb + (junc ? tb : 0); ctx->last_d2h = sizeof(KsResult) * (unsigned long long)n;
	{ const cudaError_t e = upload_seqs(0); if (e != cudaSuccess) { drain(); return ks_fail(-10, "upload failed: %s", cudaGetErrorString(e)); } }
	ksw2b_plan *pl = plan_build(ctx, par, n, qoff, toff, wv, bounds, false);
	if (!pl) { drain(); return -3; }
	const double t_plan = now();
	if (pl->prep != KS_PREP_OK) { drain(); for (int64_t i = 0; i < n; ++i) fill_reset(&res[i]); ksw2b_plan_destroy(pl); return 0; }
	if (!pl->uniform) ctx->last_h2d += sizeof(KsJob) * (unsigned long long)n;
	int rc = 0;
	// results come back through pinned staging unless the caller's buffer is itself pinned
	cudaPointerAttributes pa; bool res_pinned = (cudaPointerGetAttributes(&pa, res) == cudaSuccess && pa.type == cudaMemoryTypeHost);
	cudaGetLastError();
	if (!res_pinned && ctx->h_res.ensure(sizeof(KsResult) * (size_t)n)) { drain(); ksw2b_plan_destroy(pl); return ks_fail(-11, "pinned result staging allocation failed"); }
	ksw2b_result_t *stage = res_pinned ? res : (ksw2b_result_t*)ctx->h_res.p;
	pl->launches = 0; pl->chunk_cig_used.assign(pl->chunks.size(), 0); pl->drained = 0;
	if (pl->cig) ctx->cig_host.clear();
	ctx->last_fill_ms = ctx->last_span_ms = 0; ctx->last_fill_launches = ctx->last_launches = 0;
	if (ctx->timing) {
		pl->timing = true; pl->tev_used = 0;
		for (auto &e : ctx->tm) if (!e) CK(cudaEventCreate(&e));
	}
	// Score-only segments alternate between two compute streams (and two save arenas): the next segment's persistent CTAs move in as the
	// previous segment's run out of jobs, so a launch's tail (a warp's last 32 alignments, ~1 ms) is not dead time.  CIGAR runs share the
	// direction arena and stay on one stream.
	const bool two = !pl->cig && !pl->rows && !pl->extf && !pl->gg2 && !pl->approx && env_two != 0;   // (the row-wise kernels share per-warp scratch slots between launches)
	do {
		cudaError_t e = cudaSuccess;
		// pass 1: enqueue everything (inputs on s_in, kernels on s_cmp / s_cmp2, results out on s_out)
		for (size_t s = 0; s < pl->segs.size() && !rc; ++s) {
			const Seg &S = pl->segs[s];
			cudaStream_t sc = (two && (s & 1)) ? ctx->s_cmp2 : ctx->s_cmp;
			pl->slot = (two && (s & 1)) ? 1 : 0;
			if (s > 0 && (e = upload_seqs(s)) != cudaSuccess) break;
			if ((rc = upload_jobs(pl, S.lo, S.hi, ctx->s_in)) != 0) break;
			if ((e = cudaEventRecord(ctx->ev[4 * s], ctx->s_in)) != cudaSuccess || (e = cudaStreamWaitEvent(sc, ctx->ev[4 * s], 0)) != cudaSuccess) break;
			if (ctx->timing && s == 0 && (e = cudaEventRecord(ctx->tm[0], sc)) != cudaSuccess) break;      // (after the wait: the first segment's inputs are resident)
			for (size_t ci = S.c0; ci < S.c1 && !rc; ++ci)
				rc = run_chunk(pl, ci, (const uint8_t*)ctx->d_q.p, (const uint8_t*)ctx->d_t.p, junc ? (const uint8_t*)ctx->d_j.p : 0, sc);
			if (rc) break;
			if ((e = cudaEventRecord(ctx->ev[4 * s + 2], sc)) != cudaSuccess || (e = cudaStreamWaitEvent(ctx->s_out, ctx->ev[4 * s + 2], 0)) != cudaSuccess ||
			    (e = cudaMemcpyAsync(stage + S.lo, (KsResult*)ctx->d_res.p + S.lo, sizeof(KsResult) * (size_t)(S.hi - S.lo), cudaMemcpyDeviceToHost, ctx->s_out)) != cudaSuccess ||
			    (e = cudaEventRecord(ctx->ev[4 * s + 3], ctx->s_out)) != cudaSuccess) break;
		}
		pl->slot = 0;
		if (!rc && e != cudaSuccess) { rc = ks_fail(-10, "pipeline failed: %s", cudaGetErrorString(e)); break; }
		if (rc) break;
		const double t_enq = now();
		if (timing) fprintf(stderr, "[ksw2b_align] plan %.2f ms, enqueue %.2f ms", t_plan - t_start, t_enq - t_plan);
		// pass 2: hand results to the caller segment by segment while later segments still compute
		for (size_t s = 0; s < pl->segs.size(); ++s) {
			if ((e = cudaEventSynchronize(ctx->ev[4 * s + 3])) != cudaSuccess) { rc = ks_fail(-10, "sync failed: %s", cudaGetErrorString(e)); break; }
			if (!res_pinned) memcpy(res + pl->segs[s].lo, stage + pl->segs[s].lo, sizeof(KsResult) * (size_t)(pl->segs[s].hi - pl->segs[s].lo));
		}
		if (rc) break;
		if ((e = cudaStreamSynchronize(ctx->s_cmp)) != cudaSuccess || (e = cudaStreamSynchronize(ctx->s_cmp2)) != cudaSuccess) { rc = ks_fail(-10, "sync failed: %s", cudaGetErrorString(e)); break; }
		const double t_res = now();
		if (ctx->timing) {
			float a = 0, b = 0; int nl = 0;
			if ((e = cudaEventRecord(ctx->tm[1], ctx->s_cmp)) != cudaSuccess || (e = cudaEventRecord(ctx->tm[2], ctx->s_cmp2)) != cudaSuccess ||
			    (e = cudaEventSynchronize(ctx->tm[1])) != cudaSuccess || (e = cudaEventSynchronize(ctx->tm[2])) != cudaSuccess ||
			    (e = cudaEventElapsedTime(&a, ctx->tm[0], ctx->tm[1])) != cudaSuccess || (e = cudaEventElapsedTime(&b, ctx->tm[0], ctx->tm[2])) != cudaSuccess) {
				rc = ks_fail(-10, "timing failed: %s", cudaGetErrorString(e)); break;
			}
			ctx->last_span_ms = a > b ? a : b;
			ctx->last_fill_ms = ksw2b_plan_fill_ms(pl, &nl); ctx->last_fill_launches = nl;
		}
		ctx->last_launches = pl->launches;
		if (pl->cig) { rc = collect_cigars(pl, res, cigar, ctx->s_cmp); ctx->last_d2h += 4ull * ctx->cig_host.size(); }
		if (timing) fprintf(stderr, ", wait+results %.2f ms, cigars %.2f ms, total %.2f ms\n", t_res - t_enq, now() - t_res, now() - t_start);
	} while (0);
	if (rc) drain();
	ksw2b_plan_destroy(pl);
	return rc;
}

extern "C" int ksw2b_align(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const uint8_t *qcat, const int64_t *qoff,
                           const uint8_t *tcat, const int64_t *toff, const uint8_t *junc, ksw2b_result_t *res, const uint32_t **cigar)
{
	return ksw2b_align_ex(ctx, par, n, qcat, qoff, tcat, toff, junc, 0, res, cigar);
}

// ---- several GPUs from one caller: shard, one host thread + one context per device, results in caller order ----
struct KsDevShard {
	ksw2b_ctx *ctx = 0;
	PinBuf hq, ht, hj;
	std::vector<int64_t> idx, qoff, toff;
	std::vector<int32_t> w;
	std::vector<ksw2b_result_t> res;
	const uint32_t *cig = 0; size_t cig_words = 0;
	int rc = 0; char err[512];
	double span_ms = 0;
};
struct ksw2b_multi { std::vector<KsDevShard*> dev; std::vector<uint32_t> cig; };

extern "C" ksw2b_multi_t *ksw2b_multi_create(const int *devices, int n_dev)
{
	if (n_dev <= 0) { ks_fail(-2, "bad arguments"); return 0; }
	ksw2b_multi *m = new ksw2b_multi();
	for (int d = 0; d < n_dev; ++d) {
		KsDevShard *sh = new KsDevShard();
		sh->ctx = ksw2b_create(devices ? devices[d] : d);
		m->dev.push_back(sh);
		if (!sh->ctx) { ksw2b_multi_destroy(m); return 0; }
	}
	return m;
}
extern "C" void ksw2b_multi_destroy(ksw2b_multi_t *m)
{
	if (!m) return;
	for (KsDevShard *sh : m->dev) { if (sh->ctx) { cudaSetDevice(sh->ctx->device); sh->hq.release(); sh->ht.release(); sh->hj.release(); ksw2b_destroy(sh->ctx); } delete sh; }
	delete m;
}
extern "C" int ksw2b_multi_devices(ksw2b_multi_t *m) { return m ? (int)m->dev.size() : 0; }
extern "C" void ksw2b_multi_last(ksw2b_multi_t *m, int64_t *pairs, double *span_ms)
{
	if (!m) return;
	for (size_t d = 0; d < m->dev.size(); ++d) { if (pairs) pairs[d] = (int64_t)m->dev[d]->idx.size(); if (span_ms) span_ms[d] = m->dev[d]->span_ms; }
}

// work estimate of one pair: lanes of the band once it is full x diagonals / 2 (+ the traceback's share): SURVEY 8e
static inline int64_t pair_cost(int kind, int w, int ql, int tl, bool cigar)
{
	if (ql <= 0 || tl <= 0) return 1;
	const int we = ks_eff_w(kind, w, ql, tl), shortest = ql < tl ? ql : tl;
	const int64_t band = std::min<int64_t>(shortest, 2ll * we + 1);
	return band * ((int64_t)ql + tl) / 2 + 1 + (cigar ? ql + tl : 0);
}

extern "C" int ksw2b_multi_align(ksw2b_multi_t *m, const ksw2b_params_t *par, int64_t n, const uint8_t *qcat, const int64_t *qoff,
                                 const uint8_t *tcat, const int64_t *toff, const uint8_t *junc, const int32_t *wv, ksw2b_result_t *res, const uint32_t **cigar)
{
	if (!m || !par || !res || n < 0) return ks_fail(-2, "bad arguments");
	if (cigar) *cigar = 0;
	const int D = (int)m->dev.size();
	const bool with_cig = !(par->flag & KSF_SCORE_ONLY) && par->kind != KSW2B_EXTF2;
	for (KsDevShard *sh : m->dev) { sh->idx.clear(); sh->rc = 0; sh->cig = 0; sh->cig_words = 0; sh->span_ms = 0; }
	m->cig.clear();
	if (n == 0) return 0;
	// contiguous shards when every pair costs the same, cost-balanced otherwise
	bool uniform = true;
	for (int64_t i = 1; i < n && uniform; ++i)
		if (qoff[i + 1] - qoff[i] != qoff[1] - qoff[0] || toff[i + 1] - toff[i] != toff[1] - toff[0] || (wv && wv[i] != wv[0])) uniform = false;
	if (uniform) {
		for (int d = 0; d < D; ++d) { const int64_t lo = n * d / D, hi = n * (d + 1) / D; m->dev[d]->idx.resize((size_t)(hi - lo)); for (int64_t i = lo; i < hi; ++i) m->dev[d]->idx[(size_t)(i - lo)] = i; }
	} else {
		std::vector<std::pair<int64_t, int64_t>> order((size_t)n);
		for (int64_t i = 0; i < n; ++i)
			order[(size_t)i] = { -pair_cost(par->kind, wv ? wv[i] : par->w, (int)(qoff[i + 1] - qoff[i]), (int)(toff[i + 1] - toff[i]), with_cig), i };
		std::sort(order.begin(), order.end());
		for (int64_t k = 0; k < n; ++k) { const int pos = (int)(k % (2 * D)); m->dev[pos < D ? pos : 2 * D - 1 - pos]->idx.push_back(order[(size_t)k].second); }
		for (KsDevShard *sh : m->dev) std::sort(sh->idx.begin(), sh->idx.end());
	}
	auto work = [&](int d) {
		KsDevShard &sh = *m->dev[d];
		const int64_t k = (int64_t)sh.idx.size();
		sh.res.resize((size_t)k);
		if (k == 0) return;
		if (cudaSetDevice(sh.ctx->device) != cudaSuccess) { sh.rc = -1; snprintf(sh.err, sizeof sh.err, "cudaSetDevice(%d) failed", sh.ctx->device); return; }
		sh.qoff.resize((size_t)k + 1); sh.toff.resize((size_t)k + 1);
		const uint8_t *q = 0, *t = 0, *j = 0;
		const int32_t *w = 0;
		if (uniform) {                                     // a slice of the caller's buffers, offsets rebased
			const int64_t lo = sh.idx[0], q0 = qoff[lo], t0 = toff[lo];
			for (int64_t i = 0; i <= k; ++i) { sh.qoff[(size_t)i] = qoff[lo + i] - q0; sh.toff[(size_t)i] = toff[lo + i] - t0; }
			q = qcat + q0; t = tcat + t0; j = junc ? junc + t0 : 0; w = wv ? wv + lo : 0;
		} else {                                           // gather the shard into pinned staging (the copy doubles as the H2D source)
			sh.qoff[0] = sh.toff[0] = 0;
			for (int64_t i = 0; i < k; ++i) { const int64_t g = sh.idx[(size_t)i]; sh.qoff[(size_t)i + 1] = sh.qoff[(size_t)i] + (qoff[g + 1] - qoff[g]); sh.toff[(size_t)i + 1] = sh.toff[(size_t)i] + (toff[g + 1] - toff[g]); }
			if (sh.hq.ensure((size_t)sh.qoff[(size_t)k] + 1) || sh.ht.ensure((size_t)sh.toff[(size_t)k] + 1) || (junc && sh.hj.ensure((size_t)sh.toff[(size_t)k] + 1))) {
				sh.rc = -11; snprintf(sh.err, sizeof sh.err, "pinned staging allocation failed on device %d", sh.ctx->device); return;
			}
			if (wv) sh.w.resize((size_t)k);
			for (int64_t i = 0; i < k; ++i) {
				const int64_t g = sh.idx[(size_t)i];
				memcpy((uint8_t*)sh.hq.p + sh.qoff[(size_t)i], qcat + qoff[g], (size_t)(qoff[g + 1] - qoff[g]));
				memcpy((uint8_t*)sh.ht.p + sh.toff[(size_t)i], tcat + toff[g], (size_t)(toff[g + 1] - toff[g]));
				if (junc) memcpy((uint8_t*)sh.hj.p + sh.toff[(size_t)i], junc + toff[g], (size_t)(toff[g + 1] - toff[g]));
				if (wv) sh.w[(size_t)i] = wv[g];
			}
			q = (const uint8_t*)sh.hq.p; t = (const uint8_t*)sh.ht.p; j = junc ? (const uint8_t*)sh.hj.p : 0; w = wv ? sh.w.data() : 0;
		}
		const bool tm = sh.ctx->timing; sh.ctx->timing = true;
		sh.rc = ksw2b_align_ex(sh.ctx, par, k, q, sh.qoff.data(), t, sh.toff.data(), j, w, sh.res.data(), &sh.cig);
		sh.ctx->timing = tm;
		if (sh.rc) { snprintf(sh.err, sizeof sh.err, "device %d: %.400s", sh.ctx->device, g_err); return; }
		sh.span_ms = sh.ctx->last_span_ms;
		sh.cig_words = sh.ctx->cig_host.size();
	};
	if (D == 1) work(0);
	else { std::vector<std::thread> th; for (int d = 0; d < D; ++d) th.emplace_back(work, d); for (auto &x : th) x.join(); }
	for (KsDevShard *sh : m->dev) if (sh->rc) return ks_fail(sh->rc, "%s", sh->err);
	// results in caller order; CIGAR words of all devices in one buffer
	size_t tot = 0;
	if (with_cig) { for (KsDevShard *sh : m->dev) tot += sh->cig_words; m->cig.resize(tot); }
	size_t base = 0;
	for (KsDevShard *sh : m->dev) {
		if (with_cig && sh->cig_words) memcpy(m->cig.data() + base, sh->cig, sh->cig_words * 4);
		for (size_t i = 0; i < sh->idx.size(); ++i) { ksw2b_result_t r = sh->res[i]; if (r.n_cigar > 0) r.cigar_off += (int64_t)base; res[sh->idx[i]] = r; }
		base += sh->cig_words;
	}
	if (cigar && with_cig) *cigar = m->cig.data();
	return 0;
}

// ---- allocator bridge for ez->cigar (reference: krealloc(km, ...) in ksw_push_cigar, ksw2.h:116-119) ----
typedef void *(*krealloc_t)(void*, void*, size_t);
static krealloc_t g_krealloc = 0;
extern "C" void ksw2b_set_allocator(krealloc_t f) { g_krealloc = f; }
static void *cig_realloc(void *km, void *p, size_t sz)
{
	if (!km) return realloc(p, sz);
	if (!g_krealloc) g_krealloc = (krealloc_t)dlsym(RTLD_DEFAULT, "krealloc");
	if (!g_krealloc) { fprintf(stderr, "ksw2_b200: km != NULL but no krealloc() in the process; call ksw2b_set_allocator()\n"); abort(); }
	return g_krealloc(km, p, sz);
}

// copy one result into the caller's ksw_extz_t the way the reference leaves it (reset + fields; cigar buffer re-used, grown by doubling)
static void store_ez(void *km, const ksw2b_result_t &r, const uint32_t *cig, ksw_extz_t *ez)
{
	ez->max = (uint32_t)r.max; ez->zdropped = (uint32_t)r.zdropped; ez->max_q = r.max_q; ez->max_t = r.max_t;
	ez->mqe = r.mqe; ez->mqe_t = r.mqe_t; ez->mte = r.mte; ez->mte_q = r.mte_q; ez->score = r.score; ez->reach_end = r.reach_end;
	ez->n_cigar = 0;
	if (r.n_cigar > 0 && cig) {
		int m = ez->m_cigar;
		while (m < r.n_cigar) m = m ? m << 1 : 4;
		if (m != ez->m_cigar || !ez->cigar) { ez->cigar = (uint32_t*)cig_realloc(km, ez->cigar, (size_t)m << 2); ez->m_cigar = m; }
		memcpy(ez->cigar, cig + r.cigar_off, (size_t)r.n_cigar * 4);
		ez->n_cigar = r.n_cigar;
	}
}

// run f(t, lo, hi) over [0, n) on up to 16 host threads (large batches only)
template<class F> static void ks_parallel_for(int64_t n, int64_t grain, F f)
{
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), n / std::max<int64_t>(1, grain)));
	if (T == 1) { f(0, (int64_t)0, n); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < T; ++t) th.emplace_back(f, t, n * t / T, n * (t + 1) / T);
	for (auto &x : th) x.join();
}

static int batch_ptrs(ksw2b_ctx_t *ctx, void *km, const ksw2b_params_t *par, int64_t n, const int *qlen, const uint8_t *const *query,
                      const int *tlen, const uint8_t *const *target, const uint8_t *const *junc, ksw_extz_t *ez)
{
	if (!ctx || n < 0) return ks_fail(-2, "bad arguments");
	// gather the caller's scattered sequences into pinned staging (several host threads): the copy doubles as the H2D source
	std::vector<int64_t> qoff((size_t)n + 1, 0), toff((size_t)n + 1, 0);
	for (int64_t i = 0; i < n; ++i) { qoff[i + 1] = qoff[i] + std::max(0, qlen[i]); toff[i + 1] = toff[i] + std::max(0, tlen[i]); }
	CK(cudaSetDevice(ctx->device));
	if (ctx->h_q.ensure((size_t)qoff[n] + 1) || ctx->h_t.ensure((size_t)toff[n] + 1) || (junc && ctx->h_j.ensure((size_t)toff[n] + 1)))
		return ks_fail(-11, "pinned staging allocation failed");
	uint8_t *qcat = (uint8_t*)ctx->h_q.p, *tcat = (uint8_t*)ctx->h_t.p, *jcat = junc ? (uint8_t*)ctx->h_j.p : 0;
	ks_parallel_for(n, 32768, [&](int, int64_t lo, int64_t hi) {
		for (int64_t i = lo; i < hi; ++i) {
			if (qlen[i] > 0) memcpy(qcat + qoff[i], query[i], (size_t)qlen[i]);
			if (tlen[i] > 0) memcpy(tcat + toff[i], target[i], (size_t)tlen[i]);
			if (junc && tlen[i] > 0) { if (junc[i]) memcpy(jcat + toff[i], junc[i], (size_t)tlen[i]); else memset(jcat + toff[i], 0, (size_t)tlen[i]); }
		} });
	std::vector<ksw2b_result_t> res((size_t)n);
	const uint32_t *cig = 0;
	int rc = ksw2b_align(ctx, par, n, qcat, qoff.data(), tcat, toff.data(), jcat, res.data(), &cig);
	if (rc) return rc;
	// ez->cigar grows through the caller's allocator (km arenas are not thread-safe): CIGAR runs are stored on this thread
	if (!cig) ks_parallel_for(n, 65536, [&](int, int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) store_ez(km, res[(size_t)i], 0, &ez[i]); });
	else for (int64_t i = 0; i < n; ++i) store_ez(km, res[(size_t)i], cig, &ez[i]);
	return 0;
}

extern "C" int ksw2b_extz2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                                 const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	ksw2b_params_t p; memset(&p, 0, sizeof p);
	p.kind = KSW2B_EXTZ2; p.m = m; p.mat = mat; p.q = q; p.e = e; p.w = w; p.zdrop = zdrop; p.end_bonus = end_bonus; p.flag = flag;
	return batch_ptrs(ctx, km, &p, n, qlen, query, tlen, target, 0, ez);
}
extern "C" int ksw2b_extd2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                                 const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2,
                                 int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	ksw2b_params_t p; memset(&p, 0, sizeof p);
	p.kind = KSW2B_EXTD2; p.m = m; p.mat = mat; p.q = q; p.e = e; p.q2 = q2; p.e2 = e2; p.w = w; p.zdrop = zdrop; p.end_bonus = end_bonus; p.flag = flag;
	return batch_ptrs(ctx, km, &p, n, qlen, query, tlen, target, 0, ez);
}
extern "C" int ksw2b_exts2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                                 const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t noncan,
                                 int zdrop, int8_t junc_bonus, int flag, const uint8_t *const *junc, ksw_extz_t *ez)
{
	ksw2b_params_t p; memset(&p, 0, sizeof p);
	p.kind = KSW2B_EXTS2; p.m = m; p.mat = mat; p.q = q; p.e = e; p.q2 = q2; p.noncan = noncan; p.w = -1; p.zdrop = zdrop; p.junc_bonus = junc_bonus; p.flag = flag;
	bool any = false;
	if (junc) for (int64_t i = 0; i < n; ++i) if (junc[i]) any = true;
	return batch_ptrs(ctx, km, &p, n, qlen, query, tlen, target, any ? junc : 0, ez);
}

extern "C" int ksw2b_extz_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                                const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	ksw2b_params_t p; memset(&p, 0, sizeof p);
	p.kind = KSW2B_EXTZ; p.m = m; p.mat = mat; p.q = q; p.e = e; p.w = w; p.zdrop = zdrop; p.flag = flag;
	return batch_ptrs(ctx, km, &p, n, qlen, query, tlen, target, 0, ez);
}
extern "C" int ksw2b_extd_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                                const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2,
                                int w, int zdrop, int flag, ksw_extz_t *ez)
{
	ksw2b_params_t p; memset(&p, 0, sizeof p);
	p.kind = KSW2B_EXTD; p.m = m; p.mat = mat; p.q = q; p.e = e; p.q2 = q2; p.e2 = e2; p.w = w; p.zdrop = zdrop; p.flag = flag;
	return batch_ptrs(ctx, km, &p, n, qlen, query, tlen, target, 0, ez);
}

// ---- the unchanged single-pair entry points (reference ksw2.h:61-74) ----
// The reference API aligns one pair per call and is re-entrant: minimap2-style programs call it from many host threads at once
// (SURVEY 8b "Threading", 8f row F1).  A GPU wants batches, so concurrent calls are COMBINED: a caller queues its request; a waiting
// caller whose request is still queued and that finds a free LANE (a context of its own on the device, KS_LANES of them) leads one
// round -- takes everything queued so far, groups it by parameter set (calls that differ only in the band share a group: the band
// travels per pair), runs each group as one batch on the lane's context and wakes the others.  Calls that arrive while the lanes
// are busy form the next batches ("group commit"); several lanes let the next batch start while the previous one is still on the
// GPU, so a caller's turn-around is one batch latency (~1 ms for short pairs: one warp per pair is latency bound), not two.
// A lone caller simply runs a batch of one.  Each caller copies its own result into its own ksw_extz_t and grows ez->cigar on ITS
// thread with ITS km (kalloc arenas are per thread, kalloc.c).  A leader lingers up to 50 microseconds for company before it launches, but
// only while recent rounds did carry several calls (a lone caller never waits); KSW2B_LINGER_US=<n> fixes the wait (0: never),
// KSW2B_LANES=<1..8> sets the number of lanes (default 4).  Measured (B200, 150 bp extension pairs, scripts/combine_bench.py): 1 thread
// 1.5 k calls/s (one batch latency per call), 256 threads 137 k (183 k with the linger), 1024 threads 182 k (253 k) calls/s.
struct KsCall {
	ksw2b_params_t par;
	int qlen, tlen;
	const uint8_t *query, *target, *junc;
	ksw2b_result_t res;
	std::vector<uint32_t> cig;
	int rc = 0;
	bool done = false, queued = false;
	char err[256];
};
enum { KS_LANES_MAX = 8 };
static std::mutex g_mu;
static std::condition_variable g_cv, g_cv_arrive;
static std::vector<KsCall*> g_queue;
static ksw2b_ctx *g_lane_ctx[KS_LANES_MAX];
static bool g_lane_busy[KS_LANES_MAX];
static int g_lanes = 0;
static long g_linger_us = -1;               // < 0: adaptive (50 us while rounds carry company)
static double g_avg_round = 1.0;            // moving average of calls per round
static size_t g_max_batch = 1 << 16;
static unsigned long long g_stat_calls = 0, g_stat_batches = 0;

// same parameter block up to the band (which is carried per pair)
static bool same_params(const ksw2b_params_t &a, const ksw2b_params_t &b)
{
	const bool per_pair_band = a.kind <= KSW2B_EXTD2;
	return a.kind == b.kind && a.m == b.m && a.q == b.q && a.e == b.e && a.q2 == b.q2 && a.e2 == b.e2 && (per_pair_band || a.w == b.w) && a.zdrop == b.zdrop &&
	       a.end_bonus == b.end_bonus && a.flag == b.flag && a.noncan == b.noncan && a.junc_bonus == b.junc_bonus &&
	       (a.mat == b.mat || (a.m > 0 && a.mat && b.mat && memcmp(a.mat, b.mat, (size_t)a.m * a.m) == 0));
}

struct KsGather { std::vector<int64_t> qoff, toff; std::vector<uint8_t> qcat, tcat, jcat; std::vector<int32_t> w; std::vector<ksw2b_result_t> res; };

// one group of calls (same parameters) as one batch on ctx; fills rc / res / cig of every call
static int run_group(ksw2b_ctx *ctx, KsCall *const *grp, size_t n, KsGather &G)
{
	bool any_junc = false, same_w = true;
	for (size_t i = 0; i < n; ++i) { any_junc |= grp[i]->junc != 0; same_w &= grp[i]->par.w == grp[0]->par.w; }
	G.qoff.assign(n + 1, 0); G.toff.assign(n + 1, 0);
	for (size_t i = 0; i < n; ++i) { G.qoff[i + 1] = G.qoff[i] + std::max(0, grp[i]->qlen); G.toff[i + 1] = G.toff[i] + std::max(0, grp[i]->tlen); }
	G.qcat.resize((size_t)G.qoff[n] + 1); G.tcat.resize((size_t)G.toff[n] + 1); G.jcat.assign(any_junc ? (size_t)G.toff[n] + 1 : 0, 0);
	if (!same_w) G.w.resize(n);
	for (size_t i = 0; i < n; ++i) {
		if (grp[i]->qlen > 0) memcpy(&G.qcat[(size_t)G.qoff[i]], grp[i]->query, (size_t)grp[i]->qlen);
		if (grp[i]->tlen > 0) memcpy(&G.tcat[(size_t)G.toff[i]], grp[i]->target, (size_t)grp[i]->tlen);
		if (any_junc && grp[i]->junc && grp[i]->tlen > 0) memcpy(&G.jcat[(size_t)G.toff[i]], grp[i]->junc, (size_t)grp[i]->tlen);
		if (!same_w) G.w[i] = grp[i]->par.w;
	}
	G.res.resize(n);
	const uint32_t *cig = 0;
	const int rc = ksw2b_align_ex(ctx, &grp[0]->par, (int64_t)n, G.qcat.data(), G.qoff.data(), G.tcat.data(), G.toff.data(), any_junc ? G.jcat.data() : 0,
	                              same_w ? 0 : G.w.data(), G.res.data(), &cig);
	for (size_t i = 0; i < n; ++i) {
		KsCall &c = *grp[i];
		c.rc = rc;
		if (rc) { snprintf(c.err, sizeof c.err, "%.250s", g_err); continue; }
		c.res = G.res[i];
		if (G.res[i].n_cigar > 0 && cig) c.cig.assign(cig + G.res[i].cigar_off, cig + G.res[i].cigar_off + G.res[i].n_cigar);
		c.res.cigar_off = 0;
	}
	return rc;
}

// leader only: run every parameter group of `batch` on the lane's context
static void run_combined(std::vector<KsCall*> &batch, int lane)
{
	if (!g_lane_ctx[lane]) {
		g_lane_ctx[lane] = ksw2b_create(-1);
		if (!g_lane_ctx[lane]) { fprintf(stderr, "ksw2_b200: %s\n", g_err); abort(); }   // never fall back to a CPU path
	}
	ksw2b_ctx *ctx = g_lane_ctx[lane];
	std::vector<char> taken(batch.size(), 0);
	std::vector<KsCall*> grp;
	KsGather G;
	unsigned long long nb = 0;
	for (size_t a = 0; a < batch.size(); ++a) {
		if (taken[a]) continue;
		grp.clear();
		for (size_t b = a; b < batch.size(); ++b)
			if (!taken[b] && same_params(batch[a]->par, batch[b]->par)) { taken[b] = 1; grp.push_back(batch[b]); }
		++nb;
		// A failure of the whole group (device memory for one huge pair, shared memory of the warp mode, ...) must not take the other callers
		// down with it: the calls are retried one by one and only a call that still fails keeps its error.
		if (run_group(ctx, grp.data(), grp.size(), G) != 0 && grp.size() > 1)
			for (KsCall *c : grp) { ++nb; run_group(ctx, &c, 1, G); }
	}
	std::lock_guard<std::mutex> lk(g_mu);
	g_stat_batches += nb; g_stat_calls += batch.size();
}

static void combined_call(KsCall &c, void *km, ksw_extz_t *ez)
{
	{
		std::unique_lock<std::mutex> lk(g_mu);
		if (g_lanes == 0) {
			const char *e = getenv("KSW2B_LINGER_US"); g_linger_us = e ? atol(e) : -1;
			const char *l = getenv("KSW2B_LANES"); g_lanes = l ? atoi(l) : 4; if (g_lanes < 1) g_lanes = 1; if (g_lanes > KS_LANES_MAX) g_lanes = KS_LANES_MAX;
		}
		c.queued = true;
		g_queue.push_back(&c);
		g_cv_arrive.notify_one();
		while (!c.done) {
			int lane = -1;
			if (c.queued) for (int i = 0; i < g_lanes && lane < 0; ++i) if (!g_lane_busy[i]) lane = i;
			if (lane < 0) { g_cv.wait(lk); continue; }             // in somebody's batch, or every lane is busy
			g_lane_busy[lane] = true;                              // this caller leads one round on `lane`
			const long linger = g_linger_us >= 0 ? g_linger_us : (g_avg_round >= 1.5 ? 50 : 0);
			if (linger > 0) g_cv_arrive.wait_for(lk, std::chrono::microseconds(linger), [] { return g_queue.size() >= g_max_batch; });
			std::vector<KsCall*> batch;
			if (g_queue.size() <= g_max_batch) batch.swap(g_queue);
			else { batch.assign(g_queue.begin(), g_queue.begin() + g_max_batch); g_queue.erase(g_queue.begin(), g_queue.begin() + g_max_batch); }
			for (KsCall *b : batch) b->queued = false;
			g_avg_round = 0.75 * g_avg_round + 0.25 * (double)batch.size();
			lk.unlock();
			try { run_combined(batch, lane); }
			catch (...) { fprintf(stderr, "ksw2_b200: out of host memory while combining %zu calls\n", batch.size()); abort(); }   // (never leave the others waiting)
			lk.lock();
			for (KsCall *b : batch) b->done = true;
			g_lane_busy[lane] = false;
			g_cv.notify_all();
		}
	}
	if (c.rc) { fprintf(stderr, "ksw2_b200: alignment failed (%d): %s\n", c.rc, c.err); abort(); }
	store_ez(km, c.res, c.cig.data(), ez);
}

// statistics of the combining layer: calls served and batches launched since the library was loaded
extern "C" void ksw2b_combine_stats(unsigned long long *calls, unsigned long long *batches)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (calls) *calls = g_stat_calls;
	if (batches) *batches = g_stat_batches;
}

static void single_call(int kind, void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                        int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag, int noncan, int junc_bonus, const uint8_t *junc, ksw_extz_t *ez)
{
	KsCall c;
	memset(&c.par, 0, sizeof c.par);
	c.par.kind = kind; c.par.m = m; c.par.mat = mat; c.par.q = q; c.par.e = e; c.par.q2 = q2; c.par.e2 = e2; c.par.w = w; c.par.zdrop = zdrop;
	c.par.end_bonus = end_bonus; c.par.flag = flag; c.par.noncan = noncan; c.par.junc_bonus = junc_bonus;
	c.qlen = qlen; c.tlen = tlen; c.query = query; c.target = target; c.junc = junc;
	combined_call(c, km, ez);
}

extern "C" void ksw_extz2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                              int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTZ2, km, qlen, query, tlen, target, m, mat, q, e, 0, 0, w, zdrop, end_bonus, flag, 0, 0, 0, ez);
}
extern "C" void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                              int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTD2, km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, 0, 0, 0, ez);
}
extern "C" void ksw_exts2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                              int8_t q, int8_t e, int8_t q2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTS2, km, qlen, query, tlen, target, m, mat, q, e, q2, 0, -1, zdrop, 0, flag, noncan, junc_bonus, junc, ez);
}
extern "C" void ksw_extz(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                         int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTZ, km, qlen, query, tlen, target, m, mat, q, e, 0, 0, w, zdrop, 0, flag, 0, 0, 0, ez);
}
extern "C" void ksw_extd(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                         int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTD, km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, 0, flag, 0, 0, 0, ez);
}
extern "C" void ksw_extf2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t mch, int8_t mis, int8_t e, int w, int xdrop, ksw_extz_t *ez)
{
	single_call(KSW2B_EXTF2, km, qlen, query, tlen, target, 0, 0, mch, e, mis, 0, w, xdrop, 0, KSF_SCORE_ONLY, 0, 0, 0, ez);
}

// Global alignment entry point of ksw2.h:88 (ksw2_gg.c): score returned, CIGAR through the caller's three pointers (may all be NULL)
static int gg_call(int kind, void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                   int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	const bool with = m_cigar_ && n_cigar_ && cigar_;
	ksw_extz_t ez; memset(&ez, 0, sizeof ez);
	if (with) { ez.cigar = *cigar_; ez.m_cigar = *m_cigar_; }
	single_call(kind, km, qlen, query, tlen, target, m, mat, q, e, 0, 0, w, -1, 0, with ? 0 : KSF_SCORE_ONLY, 0, 0, 0, &ez);
	if (with) { *cigar_ = ez.cigar; *m_cigar_ = ez.m_cigar; *n_cigar_ = ez.n_cigar; }
	return ez.score;
}
extern "C" int ksw_gg(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int w,
                      int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	return gg_call(KSW2B_GG, km, qlen, query, tlen, target, m, mat, gapo, gape, w, m_cigar_, n_cigar_, cigar_);
}
extern "C" int ksw_gg2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                       int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	return gg_call(KSW2B_GG2, km, qlen, query, tlen, target, m, mat, q, e, w, m_cigar_, n_cigar_, cigar_);
}
extern "C" int ksw_gg2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                           int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	return gg_call(KSW2B_GG2_SSE, km, qlen, query, tlen, target, m, mat, q, e, w, m_cigar_, n_cigar_, cigar_);
}

// Symbol names of a KSW_CPU_DISPATCH build of the reference (ksw2_extz2_sse.c:16-24, ksw2_extd2_sse.c:25-36, ksw2_exts2_sse.c:24-35): a caller
// that was compiled against the dispatcher's per-ISA entry points links here too.  All of them are the same GPU path (SSE4.1 semantics).
#define KS_ALIAS_Z(NAME) extern "C" void NAME(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, \
		int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez) { ksw_extz2_sse(km, qlen, query, tlen, target, m, mat, q, e, w, zdrop, end_bonus, flag, ez); }
#define KS_ALIAS_D(NAME) extern "C" void NAME(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, \
		int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez) { ksw_extd2_sse(km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, ez); }
#define KS_ALIAS_S(NAME) extern "C" void NAME(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, \
		int8_t q, int8_t e, int8_t q2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez) { ksw_exts2_sse(km, qlen, query, tlen, target, m, mat, q, e, q2, noncan, zdrop, junc_bonus, flag, junc, ez); }
KS_ALIAS_Z(ksw_extz2_sse41) KS_ALIAS_Z(ksw_extz2_sse2)
KS_ALIAS_D(ksw_extd2_sse41) KS_ALIAS_D(ksw_extd2_sse2)
KS_ALIAS_S(ksw_exts2_sse41) KS_ALIAS_S(ksw_exts2_sse2)
