/* ksw2.h -- public C API of the B200-native ksw2 hot path (drop-in boundary).
 *
 * This header re-declares, with the same names, argument order, struct layout and
 * flag values, the part of lh3/ksw2's `ksw2.h` that callers link against
 * (reference: ksw2.h:6-42 constants + ksw_extz_t, ksw2.h:61-90 prototypes: every alignment entry point).  The ABI is
 * fixed by the reference (x86-64: sizeof(ksw_extz_t)==56, cigar pointer at offset 48);
 * the implementation behind it is ksw2_b200 (CUDA, sm_100a).  The reference's private
 * inline helpers (ksw_backtrack, ksw_push_cigar, ...) are NOT part of the boundary and
 * are not declared here: traceback runs on the GPU.
 */
#ifndef KSW2_B200_KSW2_H_
#define KSW2_B200_KSW2_H_

#include <stdint.h>

#define KSW_NEG_INF (-0x40000000)

/* flag bits: values fixed by reference ksw2.h:8-18 */
#define KSW_EZ_SCORE_ONLY   0x01
#define KSW_EZ_RIGHT        0x02
#define KSW_EZ_GENERIC_SC   0x04
#define KSW_EZ_APPROX_MAX   0x08
#define KSW_EZ_APPROX_DROP  0x10
#define KSW_EZ_EXTZ_ONLY    0x40
#define KSW_EZ_REV_CIGAR    0x80
#define KSW_EZ_SPLICE_FOR   0x100
#define KSW_EZ_SPLICE_REV   0x200
#define KSW_EZ_SPLICE_FLANK 0x400
#define KSW_EZ_EQX          0x800

/* BAM-style CIGAR operators used on this path (reference ksw2.h:22-27) */
#define KSW_CIGAR_MATCH  0
#define KSW_CIGAR_INS    1
#define KSW_CIGAR_DEL    2
#define KSW_CIGAR_N_SKIP 3
#define KSW_CIGAR_EQ     7
#define KSW_CIGAR_X      8

#ifdef __cplusplus
extern "C" {
#endif

/* Result record; layout identical to reference ksw2.h:33-42. */
typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t;
	int mqe, mqe_t;
	int mte, mte_q;
	int score;
	int m_cigar, n_cigar;
	int reach_end;
	uint32_t *cigar;
} ksw_extz_t;

/* Single affine gap, anti-diagonal difference recurrence.  Replaces
 * reference ksw2_extz2_sse.c:23 (prototype ksw2.h:64-65). */
void ksw_extz2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop,
                   int end_bonus, int flag, ksw_extz_t *ez);

/* Two-piece affine gap.  Replaces reference ksw2_extd2_sse.c:34 (ksw2.h:70-71). */
void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2,
                   int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);

/* Splice-aware.  Replaces reference ksw2_exts2_sse.c:33 (ksw2.h:73-74). */
void ksw_exts2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int8_t gapo2, int8_t noncan,
                   int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez);

/* The per-ISA names a KSW_CPU_DISPATCH build of the reference exports (ksw2_extz2_sse.c:16-24 etc.): same entry points. */
void ksw_extz2_sse41(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                     int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void ksw_extz2_sse2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                    int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void ksw_extd2_sse41(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                     int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void ksw_extd2_sse2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                    int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void ksw_exts2_sse41(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                     int8_t gapo, int8_t gape, int8_t gapo2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez);
void ksw_exts2_sse2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                    int8_t gapo, int8_t gape, int8_t gapo2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez);

/* Row-wise (non-rotated) formulations of the same two gap models: -inf outside the band, one maximum / Z-drop test per
 * target row, no end_bonus.  Replace reference ksw2_extz.c:6 and ksw2_extd.c:6 (prototypes ksw2.h:61-62,67-68).
 * Served by the GPU like the kernels above (one thread per pair); defined for qlen > 0, tlen > 0. */
void ksw_extz(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez);
void ksw_extd(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int flag, ksw_extz_t *ez);

/* Linear gap cost (u/v-only recurrence), X-drop on one tracked cell, score only: fills max / max_t / max_q / score / zdropped.
 * Replaces reference ksw2_extf2_sse.c:11 (prototype ksw2.h:76). */
void ksw_extf2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t mch, int8_t mis, int8_t e,
                   int w, int xdrop, ksw_extz_t *ez);

/* Global alignment, row-wise formulation: returns the score; CIGAR through the three pointers (all NULL: score only), grown with the
 * caller's allocator like ez->cigar.  Replaces reference ksw2_gg.c:6 (prototype ksw2.h:88). */
int ksw_gg(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
           int8_t gapo, int8_t gape, int w, int *m_cigar_, int *n_cigar_, uint32_t **cigar_);

/* Global alignment, anti-diagonal formulations (scalar int8 / 16-lane).  Replace reference ksw2_gg2.c:4 and ksw2_gg2_sse.c:11
 * (prototypes ksw2.h:89-90).  Unlike the reference's ksw_gg2_sse, NULL CIGAR pointers are accepted (score only). */
int ksw_gg2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
            int8_t gapo, int8_t gape, int w, int *m_cigar_, int *n_cigar_, uint32_t **cigar_);
int ksw_gg2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                int8_t gapo, int8_t gape, int w, int *m_cigar_, int *n_cigar_, uint32_t **cigar_);

#ifdef __cplusplus
}
#endif
#endif
