/* ksw2_b200.h -- additive batch API of the B200-native ksw2 hot path (C ABI, no CUDA/torch types).
 *
 * The reference API aligns ONE pair per call (ksw2.h:61-74; only caller in-tree: cli.c:74-84).  A GPU
 * needs many pairs per launch, so besides the unchanged single-pair entry points (include/ksw2.h) the
 * library exports a batch interface.  All pairs of a batch share one parameter block (the arguments of
 * ksw_extz2_sse / ksw_extd2_sse / ksw_exts2_sse other than the sequences) -- that is how minimap2-style
 * callers use ksw2.  Sequences are passed concatenated with n+1 offsets (1 byte per base, values < m).
 *
 * Two levels:
 *   ksw2b_align()                    host buffers in, host results out (H2D, kernels, D2H inside) -- the drop-in path
 *   ksw2b_plan_*()                   explicit plan object for callers that keep sequences resident in device memory
 *                                    and want to own the stream (used by bench.py for the kernel-only figure)
 * Errors: functions return 0 on success, a negative code otherwise; ksw2b_last_error() gives the text.
 * Nothing here ever computes an alignment on the CPU: without a usable CUDA device the calls fail.
 */
#ifndef KSW2_B200_H_
#define KSW2_B200_H_
#include <stddef.h>
#include <stdint.h>
#include "ksw2.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { KSW2B_EXTZ2 = 0, KSW2B_EXTD2 = 1, KSW2B_EXTS2 = 2,
       KSW2B_EXTZ = 3, KSW2B_EXTD = 4,     /* the row-wise entry points ksw_extz / ksw_extd (reference ksw2.h:61-62,67-68) */
       KSW2B_EXTF2 = 5,                  /* ksw_extf2_sse (ksw2.h:76): q = mch, q2 = mis, e = e, zdrop = xdrop; m / mat unused; score only */
       KSW2B_GG = 6,                       /* ksw_gg (ksw2.h:88): global, row-wise; result in `score`, CIGAR unless KSW_EZ_SCORE_ONLY */
       KSW2B_GG2 = 7, KSW2B_GG2_SSE = 8 }; /* ksw_gg2 / ksw_gg2_sse (ksw2.h:89-90): global, anti-diagonal; same reporting as KSW2B_GG */

typedef struct {
	int kind;                 /* KSW2B_EXTZ2 / EXTD2 / EXTS2 / EXTZ / EXTD: which reference entry point's semantics */
	int m;                    /* alphabet size; code m-1 is the wildcard unless KSW_EZ_GENERIC_SC */
	const int8_t *mat;        /* m*m scores (host pointer) */
	int q, e, q2, e2;         /* gap open/extend; extd2: second piece; exts2: q2 = long-gap open, e2 unused */
	int w, zdrop, end_bonus;  /* band (<0 none; ignored by exts2), Z-drop (<0 off), end bonus (extz2/extd2) */
	int flag;                 /* KSW_EZ_* */
	int noncan, junc_bonus;   /* exts2 only */
} ksw2b_params_t;

/* One result per pair: the scalar fields of ksw_extz_t (ksw2.h:33-42) + where the CIGAR is. */
typedef struct {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar;
	int32_t tb_i, tb_j;       /* traceback start cell (diagnostic) */
	int32_t n_diag;           /* anti-diagonals the reference evaluates for this pair (for exact cell counts) */
	int64_t cigar_off;        /* word offset into the CIGAR buffer returned next to the results */
} ksw2b_result_t;

typedef struct ksw2b_ctx ksw2b_ctx_t;
typedef struct ksw2b_plan ksw2b_plan_t;

/* Context = one CUDA device + reusable device/pinned buffers.  device < 0: the current device. Not thread-safe:
 * use one context per host thread (like one kalloc arena per thread in the reference, kalloc.c).
 * A plan BORROWS its context's buffers (job table, coded sequences, scratch, direction arena, scoring matrix), so a context
 * carries at most ONE live plan: ksw2b_plan_create() and ksw2b_align() (which plans internally) fail with code -4 while
 * another plan of the same context is alive.  Destroy plans before their context. */
ksw2b_ctx_t *ksw2b_create(int device);
void ksw2b_destroy(ksw2b_ctx_t *ctx);
const char *ksw2b_last_error(void);

/* Drop-in batch call.  qoff/toff have n+1 entries; junc (exts2; may be NULL) is indexed like the target.
 * res[n] is filled; *cigar points to a library-owned host buffer (valid until the next call on ctx). */
int ksw2b_align(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n,
                const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *junc,
                ksw2b_result_t *res, const uint32_t **cigar);

/* Same with a band per pair (w[i] replaces par->w for pair i; w == NULL: par->w for all).  minimap2-style callers compute the band
 * per call (reference argument `w`, ksw2.h:64-71), so a batch collected from such calls carries one band per pair.
 * Only for the anti-diagonal kinds KSW2B_EXTZ2 / EXTD2 (exts2 has no band). */
int ksw2b_align_ex(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n,
                   const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *junc,
                   const int32_t *w, ksw2b_result_t *res, const uint32_t **cigar);

/* Bytes the last ksw2b_align() on ctx moved host->device (sequences, junctions, job table unless it was generated on the device)
 * and device->host (result records, CIGAR words). */
void ksw2b_last_transfer_bytes(ksw2b_ctx_t *ctx, unsigned long long *h2d, unsigned long long *d2h);

/* Optional device timing of ksw2b_align(): with timing on, the call brackets its kernels with CUDA events on the launching streams.
 * ksw2b_last_timing() then reports, for the last call on ctx: the summed device time of its DP-fill launches and their number, the
 * device span from the first kernel (first segment's inputs resident) to the end of the last kernel, and the kernels launched. */
void ksw2b_set_timing(ksw2b_ctx_t *ctx, int on);
void ksw2b_last_timing(ksw2b_ctx_t *ctx, double *fill_ms, int *fill_launches, double *span_ms, int *launches);

/* ---- several GPUs of one box from ONE caller ----
 * The pairs of a batch are independent: ksw2b_multi_align() shards them over the devices of the set -- contiguously when all pairs have
 * the same lengths, otherwise cost-balanced (cost = band cells, + qlen + tlen with a CIGAR; largest first, dealt in serpentine order so
 * every GPU gets the same mix of lengths) -- runs one host thread and one context per device, and writes the results in the caller's
 * order.  No collective and no peer traffic is involved.  res[i].cigar_off indexes the buffer returned in *cigar (library-owned, valid
 * until the next call on the set).  devices == NULL: devices 0 .. n_dev-1. */
typedef struct ksw2b_multi ksw2b_multi_t;
ksw2b_multi_t *ksw2b_multi_create(const int *devices, int n_dev);
void ksw2b_multi_destroy(ksw2b_multi_t *set);
int ksw2b_multi_devices(ksw2b_multi_t *set);
int ksw2b_multi_align(ksw2b_multi_t *set, const ksw2b_params_t *par, int64_t n,
                      const uint8_t *qcat, const int64_t *qoff, const uint8_t *tcat, const int64_t *toff, const uint8_t *junc,
                      const int32_t *w, ksw2b_result_t *res, const uint32_t **cigar);
/* pairs per device and device span (ms, see ksw2b_last_timing) of the last ksw2b_multi_align; arrays of ksw2b_multi_devices() entries */
void ksw2b_multi_last(ksw2b_multi_t *set, int64_t *pairs, double *span_ms);

/* Array-of-pointers flavour mirroring the reference argument lists; fills ez[i] exactly like n single calls would
 * (ez[i].cigar grown with the caller's allocator, see ksw2b_set_allocator). */
int ksw2b_extz2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                      const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop,
                      int end_bonus, int flag, ksw_extz_t *ez);
int ksw2b_extd2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                      const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2,
                      int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
int ksw2b_exts2_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                      const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t noncan,
                      int zdrop, int8_t junc_bonus, int flag, const uint8_t *const *junc, ksw_extz_t *ez);

/* the row-wise entry points ksw_extz / ksw_extd as batches (no end_bonus, like the reference prototypes) */
int ksw2b_extz_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                     const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez);
int ksw2b_extd_batch(ksw2b_ctx_t *ctx, void *km, int64_t n, const int *qlen, const uint8_t *const *query, const int *tlen,
                     const uint8_t *const *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2,
                     int w, int zdrop, int flag, ksw_extz_t *ez);

/* ez->cigar must be (re)allocated with the CALLER's allocator (reference: krealloc(km, ..) in ksw2.h:116-119).
 * Default: km == NULL -> libc realloc; km != NULL -> symbol `krealloc` looked up in the process (the caller's kalloc).
 * A caller can also install it explicitly. */
void ksw2b_set_allocator(void *(*krealloc_fn)(void *km, void *ptr, size_t size));

/* The single-pair entry points of ksw2.h combine concurrent calls from many host threads into GPU batches (group commit on a
 * process-wide context).  Counters since load: calls served, batches launched. */
void ksw2b_combine_stats(unsigned long long *calls, unsigned long long *batches);

/* ---- plan API (device-resident inputs) ---- */
/* Builds the per-pair job table for n pairs of the given lengths, uploads it and sizes all scratch.  `stream` is a
 * cudaStream_t passed as void* (NULL: default stream). */
ksw2b_plan_t *ksw2b_plan_create(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const int64_t *qoff, const int64_t *toff);
ksw2b_plan_t *ksw2b_plan_create_ex(ksw2b_ctx_t *ctx, const ksw2b_params_t *par, int64_t n, const int64_t *qoff, const int64_t *toff,
                                   const int32_t *w);   /* w: band per pair or NULL (see ksw2b_align_ex) */
/* Launches fill (+ traceback) for all pairs; d_* are DEVICE pointers to the concatenated sequences. Asynchronous. */
int ksw2b_plan_run(ksw2b_plan_t *plan, const uint8_t *d_qcat, const uint8_t *d_tcat, const uint8_t *d_junc, void *stream);
/* Copies results (and CIGARs) to the host; synchronises the stream. */
int ksw2b_plan_fetch(ksw2b_plan_t *plan, ksw2b_result_t *res, const uint32_t **cigar, void *stream);
const ksw2b_result_t *ksw2b_plan_device_results(ksw2b_plan_t *plan);   /* device pointer, n records */
int64_t ksw2b_plan_cells(ksw2b_plan_t *plan);                          /* in-band cells if no early exit */
int ksw2b_plan_launches(ksw2b_plan_t *plan);                           /* kernels launched by the last run */
/* Optional per-kernel timing: with timing on, ksw2b_plan_run() brackets every DP-fill launch with CUDA events on the
 * launching stream; ksw2b_plan_fill_ms() waits for them and returns their summed device time in ms (and how many
 * launches that was).  Used by bench.py for the roofline of the dominant kernel. */
void ksw2b_plan_set_timing(ksw2b_plan_t *plan, int on);
double ksw2b_plan_fill_ms(ksw2b_plan_t *plan, int *n_launches);
void ksw2b_plan_destroy(ksw2b_plan_t *plan);

/* pinned host memory helpers (so callers can stage without an extra copy) */
void *ksw2b_host_alloc(size_t bytes);
void ksw2b_host_free(void *p);

/* tuning knobs (optional): panel height C (diagonals per sweep), threads per CTA, CTAs per SM; 0 keeps the default */
void ksw2b_set_tuning(ksw2b_ctx_t *ctx, int panel, int threads, int ctas_per_sm);
/* work decomposition: 0 = automatic (decided per launch), 1 = one thread per alignment (many pairs), 2 = one warp per alignment (few long
 * pairs), 3 = one CTA per alignment (a handful of very long pairs: latency), 4 = one warp per alignment on the ring schedule (banded
 * pairs, effective band <= 512; other pairs keep the automatic choice);
 * warp_panel: diagonals per sweep of the warp mode (0 keeps the default) */
void ksw2b_set_mode(ksw2b_ctx_t *ctx, int mode, int warp_panel);

#ifdef __cplusplus
}
#endif
#endif
