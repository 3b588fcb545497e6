/* ksw2_oracle.h -- TEST INFRASTRUCTURE ONLY. Prototypes of the CPU restatement (oracle/ksw2_oracle.c).
 * Same argument lists as the reference entry points (ksw2.h:64-74); `km` is ignored (libc malloc). */
#ifndef KSW2_ORACLE_H_
#define KSW2_ORACLE_H_
#include <stdint.h>
#include "../include/ksw2.h"
#ifdef __cplusplus
extern "C" {
#endif
void kso_extz2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void kso_extd2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void kso_exts2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int8_t q2, int8_t noncan, int zdrop, int8_t junc_bonus, int flag, const uint8_t *junc, ksw_extz_t *ez);
/* row-wise scalar entry points (reference ksw2.h:61-62; ksw2_extz.c, ksw2_extd.c) */
void kso_extz(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t gapo, int8_t gape, int w, int zdrop, int flag, ksw_extz_t *ez);
void kso_extd(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
              int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int flag, ksw_extz_t *ez);
void kso_extf2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t mch, int8_t mis, int8_t e, int w, int xdrop, ksw_extz_t *ez);
int kso_gg(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int w,
           int *m_cigar_, int *n_cigar_, uint32_t **cigar_);
int kso_gg2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
            int *m_cigar_, int *n_cigar_, uint32_t **cigar_);
int kso_gg2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat, int8_t q, int8_t e, int w,
                int *m_cigar_, int *n_cigar_, uint32_t **cigar_);
int64_t kso_last_cells(void); /* in-band cells evaluated by the last call made on the calling thread */
#ifdef __cplusplus
}
#endif
#endif
