/* ksw2b_gen.c -- synthetic workload generators for bench.py and the parity tests, plain host C.
 *
 * Every pair is a pure function of (model, seed, pair index): a rank of a multi-GPU run can generate exactly the pairs of its
 * shard, the CPU baseline can generate the same sample, and the result does not depend on the number of threads.
 *
 *   model 3  (BASELINE config 3)  target L random ACGT; query = ONT-like copy: 3 % substitutions, 3.5 % insertions, 3.5 % deletions,
 *                                 indel lengths geometric(p = 0.7)
 *   model 4  (BASELINE config 4)  target L random ACGT; query = ~90 %-identity copy: 7 % substitutions, 1.5 % + 1.5 % single-base indels
 *   model 5  (BASELINE config 5)  target length log-uniform in [150, 20000]; divergence d uniform in [1 %, 12 %] split 60/20/20 into
 *                                 substitutions / insertions / deletions, indel lengths geometric(0.7)
 * Not part of the product library (nothing in ksw2_b200/ links it).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t splitmix64(uint64_t *x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rnext(rng_t *r)      /* xoshiro256** */
{
	uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
	s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
	return res;
}
static void rseed(rng_t *r, uint64_t seed, uint64_t idx)
{
	uint64_t x = seed * 0xD1342543DE82EF95ull + idx * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}
static inline double runif(rng_t *r) { return (double)(rnext(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int rgeo(rng_t *r, double p) { int k = 1; while (runif(r) >= p && k < 64) ++k; return k; }

static int pair_tlen(int model, rng_t *r, int L)
{
	if (model != 5) return L;
	const double lo = log(150.0), hi = log(20000.0);
	int l = (int)exp(lo + (hi - lo) * runif(r));
	return l < 150 ? 150 : l > 20000 ? 20000 : l;
}
static int64_t qslot_bytes(int tlen) { return (int64_t)tlen + tlen / 3 + 256; }

/* target lengths of pairs idx[0..m) (idx == NULL: 0..m-1) */
void ksg_lengths(int model, uint64_t seed, int L, int64_t m, const int64_t *idx, int32_t *tlen)
{
	for (int64_t k = 0; k < m; ++k) { rng_t r; rseed(&r, seed, (uint64_t)(idx ? idx[k] : k)); tlen[k] = pair_tlen(model, &r, L); }
}

static int gen_pair(int model, uint64_t seed, int64_t i, int L, uint8_t *t, uint8_t *q, int64_t qcap)
{
	rng_t r; rseed(&r, seed, (uint64_t)i);
	const int tl = pair_tlen(model, &r, L);
	double sub, ins, del, geo;
	if (model == 3) { sub = 0.03; ins = 0.035; del = 0.035; geo = 0.7; }
	else if (model == 4) { sub = 0.07; ins = 0.015; del = 0.015; geo = 1.0; }
	else { const double d = 0.01 + 0.11 * runif(&r); sub = 0.6 * d; ins = 0.2 * d; del = 0.2 * d; geo = 0.7; }
	uint64_t bits = 0; int nb = 0;
	for (int k = 0; k < tl; ++k) { if (nb == 0) { bits = rnext(&r); nb = 32; } t[k] = (uint8_t)(bits & 3); bits >>= 2; --nb; }
	int64_t ql = 0; int skip = 0;
	for (int k = 0; k < tl; ++k) {
		if (skip > 0) { --skip; continue; }
		const double u = runif(&r);
		if (u < del) { skip = (geo < 1.0 ? rgeo(&r, geo) : 1) - 1; continue; }
		if (u < del + ins) { int n = geo < 1.0 ? rgeo(&r, geo) : 1; while (n-- > 0 && ql < qcap) q[ql++] = (uint8_t)(rnext(&r) & 3); }
		if (ql >= qcap) break;
		q[ql++] = u > 1.0 - sub ? (uint8_t)((t[k] + 1 + rnext(&r) % 3) & 3) : t[k];
	}
	if (ql == 0) q[ql++] = t[0];
	return (int)ql;
}

typedef struct { int model, L; uint64_t seed; int64_t m; const int64_t *idx; uint8_t *tcat, *qslots; const int64_t *toff, *soff; int32_t *qlen; int tid, nthr; } job_t;
static void *worker(void *p)
{
	job_t *j = (job_t*)p;
	for (int64_t k = j->tid; k < j->m; k += j->nthr)
		j->qlen[k] = gen_pair(j->model, j->seed, j->idx ? j->idx[k] : k, j->L, j->tcat + j->toff[k], j->qslots + j->soff[k], j->soff[k + 1] - j->soff[k]);
	return 0;
}

/* bytes the caller must provide: *tbytes for tcat (exact), *qbytes for qcat (upper bound) */
void ksg_sizes(int model, uint64_t seed, int L, int64_t m, const int64_t *idx, int64_t *tbytes, int64_t *qbytes)
{
	int64_t a = 0, b = 0;
	for (int64_t k = 0; k < m; ++k) { rng_t r; rseed(&r, seed, (uint64_t)(idx ? idx[k] : k)); const int tl = pair_tlen(model, &r, L); a += tl; b += qslot_bytes(tl); }
	*tbytes = a; *qbytes = b;
}

/* Generates pairs idx[0..m) (idx == NULL: 0..m-1).  toff/qoff get m+1 offsets; queries are written into per-pair slots of qcat and then
 * compacted in place.  Returns the total query bytes (<0 on error). */
int64_t ksg_generate(int model, uint64_t seed, int L, int64_t m, const int64_t *idx, int nthreads, uint8_t *tcat, int64_t *toff, uint8_t *qcat, int64_t *qoff)
{
	if (m <= 0) { toff[0] = qoff[0] = 0; return 0; }
	int64_t *soff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + 1));
	int32_t *qlen = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
	if (!soff || !qlen) { free(soff); free(qlen); return -1; }
	toff[0] = 0; soff[0] = 0;
	for (int64_t k = 0; k < m; ++k) { rng_t r; rseed(&r, seed, (uint64_t)(idx ? idx[k] : k)); const int tl = pair_tlen(model, &r, L); toff[k + 1] = toff[k] + tl; soff[k + 1] = soff[k] + qslot_bytes(tl); }
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	if ((int64_t)nthreads > m) nthreads = (int)m;
	pthread_t th[256]; job_t jb[256];
	for (int t = 0; t < nthreads; ++t) {
		job_t j = { model, L, seed, m, idx, tcat, qcat, toff, soff, qlen, t, nthreads };
		jb[t] = j;
		if (pthread_create(&th[t], 0, worker, &jb[t]) != 0) { for (int u = 0; u < t; ++u) pthread_join(th[u], 0); free(soff); free(qlen); return -2; }
	}
	for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
	qoff[0] = 0;
	for (int64_t k = 0; k < m; ++k) { memmove(qcat + qoff[k], qcat + soff[k], (size_t)qlen[k]); qoff[k + 1] = qoff[k] + qlen[k]; }
	const int64_t tot = qoff[m];
	free(soff); free(qlen);
	return tot;
}

/* Cell accounting: per pair, in-band cells = sum over the first ndiag[i] diagonals of en0 - st0 + 1 with the reference's
 * band geometry (ksw2_extz2_sse.c:105-116), and the padded direction-byte lanes (16-rounded, :92,195) + 8 bytes per diagonal.
 * w[i] < 0: no band.  ndiag == NULL: all qlen + tlen - 1 diagonals. */
typedef struct { int64_t n; const int32_t *qlen, *tlen, *w, *ndiag; int64_t *cells, *lanes; int tid, nthr; } cjob_t;
static void *cells_worker(void *p)
{
	cjob_t *j = (cjob_t*)p;
	for (int64_t i = j->tid; i < j->n; i += j->nthr) {
		const int ql = j->qlen[i], tl = j->tlen[i];
		int64_t c = 0, l = 0;
		if (ql > 0 && tl > 0) {
			const int mx = ql > tl ? ql : tl, w = (j->w[i] < 0 || j->w[i] > mx) ? mx : j->w[i];
			int nd = ql + tl - 1;
			if (j->ndiag && j->ndiag[i] < nd) nd = j->ndiag[i];
			for (int r = 0; r < nd; ++r) {
				int st0 = 0, en0 = tl - 1;
				if (st0 < r - ql + 1) st0 = r - ql + 1;
				if (st0 < ((r - w + 1) >> 1)) st0 = (r - w + 1) >> 1;
				if (en0 > r) en0 = r;
				if (en0 > ((r + w) >> 1)) en0 = (r + w) >> 1;
				if (st0 > en0) break;
				c += en0 - st0 + 1; l += ((en0 | 15) - (st0 & ~15) + 1) + 8;
			}
		}
		if (j->cells) j->cells[i] = c;
		if (j->lanes) j->lanes[i] = l;
	}
	return 0;
}
void ksg_cells(int64_t n, const int32_t *qlen, const int32_t *tlen, const int32_t *w, const int32_t *ndiag, int64_t *cells, int64_t *lanes, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	pthread_t th[256]; cjob_t jb[256];
	int started = 0;
	for (int t = 0; t < nthreads; ++t) {
		cjob_t j = { n, qlen, tlen, w, ndiag, cells, lanes, t, nthreads };
		jb[t] = j;
		if (pthread_create(&th[t], 0, cells_worker, &jb[t]) != 0) break;
		++started;
	}
	for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
	if (started < nthreads) { cjob_t j = { n, qlen, tlen, w, ndiag, cells, lanes, 0, 1 }; cells_worker(&j); }   /* (could not start the threads: do it all here) */
}
