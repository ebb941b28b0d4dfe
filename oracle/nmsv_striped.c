/*
 * nmsv_striped.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * "parasail-equivalent" CPU engine: a from-memory restatement of the code path parasail.ssw() takes on an AVX2
 * host (upstream parasail src/ssw.c -> sw_striped_profile_avx2_256_8, re-run in _16 when the 8-bit pass
 * saturates, then the same on the reversed prefixes to find the begin cell).  Farrar striped layout, lazy-F
 * loop, end column = first column whose maximum strictly raises the running maximum, end row = smallest row in
 * the saved copy of that column (SSW-library rule).  It is used for two things only:
 *   (1) an independent second formulation the scalar oracle (nmsv_oracle.c) is cross-checked against;
 *   (2) the timed host-CPU baseline of bench.py (cpu_baseline / --impl reference), multi-threaded with pthreads,
 *       one alignment per thread at a time (mirrors nanomonsv --processes sharding, reference run.py:324-379).
 * Like Farrar's algorithm (and parasail's code) the lazy-F early exit is exact only for gap_open > gap_extend;
 * the validation path always calls with open=3, extend=1 (count_sread_by_alignment.py:148).
 * It is NOT parasail and is reported as "parasail-equivalent (port)".  PARITY UNPINNED against real parasail.
 *
 * Reference call sites this stands in for: nanomonsv/count_sread_by_alignment.py:148-149.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    int32_t score1, ref_begin1, ref_end1, read_begin1, read_end1;
} ora_ssw_result;

#define NCODE 5 /* A C G T wildcard */

static inline int sub_score(int a, int b, int match, int mismatch)
{
    if (a > 3 || b > 3) return 0;
    return a == b ? match : mismatch;
}

static void *aligned_malloc32(size_t bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 32, bytes ? bytes : 32) != 0) return NULL;
    return p;
}

/* shift a 256-bit vector left by one 16-bit / 8-bit element, filling with zero */
static inline __m256i shl_epi16_1(__m256i a)
{
    __m256i t = _mm256_permute2x128_si256(a, a, 0x08);
    return _mm256_alignr_epi8(a, t, 14);
}
static inline __m256i shl_epi8_1(__m256i a)
{
    __m256i t = _mm256_permute2x128_si256(a, a, 0x08);
    return _mm256_alignr_epi8(a, t, 15);
}

typedef struct {
    int saturated;
    int score, end_query, end_ref;
} pass_result;

/* sequence accessor: code of element i of a (possibly reversed) view */
#define SEQ(c, o, d, i) ((c)[(o) + (int64_t)(d) * (i)])

/* ---------------------------------------------------------------- 16-bit pass */
static int sw_striped_16(const uint8_t *c1, int64_t o1, int d1, int m,
                         const uint8_t *c2, int64_t o2, int d2, int n,
                         int match, int mismatch, int open, int ext, pass_result *pr)
{
    const int W = 16;
    const int segLen = (m + W - 1) / W;
    const int16_t NEG = INT16_MIN / 2;
    const int pos_limit = INT16_MAX - match - 1;
    size_t vecs = (size_t)(segLen > 0 ? segLen : 1);
    __m256i *prof = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs * NCODE);
    __m256i *pvHStore = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvHLoad = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvHMax = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvE = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    if (!prof || !pvHStore || !pvHLoad || !pvHMax || !pvE) {
        free(prof); free(pvHStore); free(pvHLoad); free(pvHMax); free(pvE);
        return -1;
    }
    /* striped query profile: element (segment i, lane l) <-> row i + l*segLen; padding rows score 0 */
    for (int b = 0; b < NCODE; ++b) {
        int16_t *t = (int16_t *)(prof + (size_t)b * vecs);
        for (int i = 0; i < segLen; ++i)
            for (int l = 0; l < W; ++l) {
                int row = i + l * segLen;
                t[i * W + l] = row < m ? (int16_t)sub_score(SEQ(c1, o1, d1, row), b, match, mismatch) : 0;
            }
    }
    const __m256i vZero = _mm256_setzero_si256();
    const __m256i vGapO = _mm256_set1_epi16((int16_t)open);
    const __m256i vGapE = _mm256_set1_epi16((int16_t)ext);
    for (int i = 0; i < segLen; ++i) {
        pvHStore[i] = vZero; pvHMax[i] = vZero;
        pvE[i] = _mm256_set1_epi16((int16_t)-open);
    }
    __m256i vMaxH = vZero, vMaxHUnit = vZero;
    int score = 0, end_ref = 0, end_query = m - 1, saturated = 0;

    for (int j = 0; j < n; ++j) {
        const __m256i *vP = prof + (size_t)SEQ(c2, o2, d2, j) * vecs;
        __m256i vF = _mm256_set1_epi16(NEG);
        __m256i vH = shl_epi16_1(pvHStore[segLen - 1]);
        __m256i *tmp = pvHLoad; pvHLoad = pvHStore; pvHStore = tmp;
        for (int i = 0; i < segLen; ++i) {
            vH = _mm256_adds_epi16(vH, vP[i]);
            __m256i vE = pvE[i];
            vH = _mm256_max_epi16(vH, vE);
            vH = _mm256_max_epi16(vH, vF);
            vH = _mm256_max_epi16(vH, vZero);
            pvHStore[i] = vH;
            vMaxH = _mm256_max_epi16(vMaxH, vH);
            vH = _mm256_subs_epi16(vH, vGapO);
            vE = _mm256_subs_epi16(vE, vGapE);
            vE = _mm256_max_epi16(vE, vH);
            pvE[i] = vE;
            vF = _mm256_subs_epi16(vF, vGapE);
            vF = _mm256_max_epi16(vF, vH);
            vH = pvHLoad[i];
        }
        /* lazy-F loop */
        for (int k = 0; k < W; ++k) {
            vF = shl_epi16_1(vF);
            vF = _mm256_insert_epi16(vF, NEG, 0);
            for (int i = 0; i < segLen; ++i) {
                vH = pvHStore[i];
                vH = _mm256_max_epi16(vH, vF);
                pvHStore[i] = vH;
                vMaxH = _mm256_max_epi16(vMaxH, vH);
                vH = _mm256_subs_epi16(vH, vGapO);
                vF = _mm256_subs_epi16(vF, vGapE);
                if (!_mm256_movemask_epi8(_mm256_cmpgt_epi16(vF, vH))) goto lazy_done;
            }
        }
lazy_done:
        if (_mm256_movemask_epi8(_mm256_cmpgt_epi16(vMaxH, vMaxHUnit))) {
            int16_t lanes[16];
            _mm256_storeu_si256((__m256i *)lanes, vMaxH);
            int mx = lanes[0];
            for (int l = 1; l < W; ++l) if (lanes[l] > mx) mx = lanes[l];
            score = mx;
            if (score > pos_limit) { saturated = 1; break; }
            vMaxHUnit = _mm256_set1_epi16((int16_t)score);
            vMaxH = vMaxHUnit;
            end_ref = j;
            memcpy(pvHMax, pvHStore, sizeof(__m256i) * vecs);
        }
    }
    if (!saturated && score > 0) {
        const int16_t *t = (const int16_t *)pvHMax;
        int column_len = segLen * W;
        end_query = m - 1;
        for (int i = 0; i < column_len; ++i) {
            if (t[i] == score) {
                int row = i / W + (i % W) * segLen;
                if (row < end_query) end_query = row;
            }
        }
    }
    pr->saturated = saturated; pr->score = score; pr->end_query = end_query; pr->end_ref = end_ref;
    free(prof); free(pvHStore); free(pvHLoad); free(pvHMax); free(pvE);
    return 0;
}

/* ---------------------------------------------------------------- 8-bit pass (biased by INT8_MIN) */
static int sw_striped_8(const uint8_t *c1, int64_t o1, int d1, int m,
                        const uint8_t *c2, int64_t o2, int d2, int n,
                        int match, int mismatch, int open, int ext, pass_result *pr)
{
    const int W = 32;
    const int segLen = (m + W - 1) / W;
    const int bias = INT8_MIN;
    const int pos_limit = INT8_MAX - match - 1; /* in biased units */
    size_t vecs = (size_t)(segLen > 0 ? segLen : 1);
    __m256i *prof = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs * NCODE);
    __m256i *pvHStore = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvHLoad = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvHMax = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    __m256i *pvE = (__m256i *)aligned_malloc32(sizeof(__m256i) * vecs);
    if (!prof || !pvHStore || !pvHLoad || !pvHMax || !pvE) {
        free(prof); free(pvHStore); free(pvHLoad); free(pvHMax); free(pvE);
        return -1;
    }
    for (int b = 0; b < NCODE; ++b) {
        int8_t *t = (int8_t *)(prof + (size_t)b * vecs);
        for (int i = 0; i < segLen; ++i)
            for (int l = 0; l < W; ++l) {
                int row = i + l * segLen;
                t[i * W + l] = row < m ? (int8_t)sub_score(SEQ(c1, o1, d1, row), b, match, mismatch) : 0;
            }
    }
    const __m256i vZero = _mm256_set1_epi8((int8_t)bias); /* biased zero; saturating adds floor H at 0 */
    const __m256i vGapO = _mm256_set1_epi8((int8_t)open);
    const __m256i vGapE = _mm256_set1_epi8((int8_t)ext);
    for (int i = 0; i < segLen; ++i) { pvHStore[i] = vZero; pvHMax[i] = vZero; pvE[i] = vZero; }
    __m256i vMaxH = vZero, vMaxHUnit = vZero;
    int score = bias, end_ref = 0, end_query = m - 1, saturated = 0;

    for (int j = 0; j < n; ++j) {
        const __m256i *vP = prof + (size_t)SEQ(c2, o2, d2, j) * vecs;
        __m256i vF = vZero;
        __m256i vH = shl_epi8_1(pvHStore[segLen - 1]);
        vH = _mm256_insert_epi8(vH, (int8_t)bias, 0);
        __m256i *tmp = pvHLoad; pvHLoad = pvHStore; pvHStore = tmp;
        for (int i = 0; i < segLen; ++i) {
            vH = _mm256_adds_epi8(vH, vP[i]);
            __m256i vE = pvE[i];
            vH = _mm256_max_epi8(vH, vE);
            vH = _mm256_max_epi8(vH, vF);
            pvHStore[i] = vH;
            vMaxH = _mm256_max_epi8(vMaxH, vH);
            vH = _mm256_subs_epi8(vH, vGapO);
            vE = _mm256_subs_epi8(vE, vGapE);
            vE = _mm256_max_epi8(vE, vH);
            pvE[i] = vE;
            vF = _mm256_subs_epi8(vF, vGapE);
            vF = _mm256_max_epi8(vF, vH);
            vH = pvHLoad[i];
        }
        for (int k = 0; k < W; ++k) {
            vF = shl_epi8_1(vF);
            vF = _mm256_insert_epi8(vF, (int8_t)bias, 0);
            for (int i = 0; i < segLen; ++i) {
                vH = pvHStore[i];
                vH = _mm256_max_epi8(vH, vF);
                pvHStore[i] = vH;
                vMaxH = _mm256_max_epi8(vMaxH, vH);
                vH = _mm256_subs_epi8(vH, vGapO);
                vF = _mm256_subs_epi8(vF, vGapE);
                if (!_mm256_movemask_epi8(_mm256_cmpgt_epi8(vF, vH))) goto lazy_done8;
            }
        }
lazy_done8:
        if (_mm256_movemask_epi8(_mm256_cmpgt_epi8(vMaxH, vMaxHUnit))) {
            int8_t lanes[32];
            _mm256_storeu_si256((__m256i *)lanes, vMaxH);
            int mx = lanes[0];
            for (int l = 1; l < W; ++l) if (lanes[l] > mx) mx = lanes[l];
            score = mx;
            if (score > pos_limit) { saturated = 1; break; }
            vMaxHUnit = _mm256_set1_epi8((int8_t)score);
            vMaxH = vMaxHUnit;
            end_ref = j;
            memcpy(pvHMax, pvHStore, sizeof(__m256i) * vecs);
        }
    }
    if (!saturated && score > bias) {
        const int8_t *t = (const int8_t *)pvHMax;
        int column_len = segLen * W;
        end_query = m - 1;
        for (int i = 0; i < column_len; ++i) {
            if (t[i] == score) {
                int row = i / W + (i % W) * segLen;
                if (row < end_query) end_query = row;
            }
        }
    }
    pr->saturated = saturated; pr->score = score - bias; pr->end_query = end_query; pr->end_ref = end_ref;
    free(prof); free(pvHStore); free(pvHLoad); free(pvHMax); free(pvE);
    return 0;
}

/* 8-bit first, 16-bit when saturated (parasail src/ssw.c policy). Returns 0 ok, 1 saturated in 16 bit, -1 error. */
static int sw_striped_sat(const uint8_t *c1, int64_t o1, int d1, int m,
                          const uint8_t *c2, int64_t o2, int d2, int n,
                          int match, int mismatch, int open, int ext, int force16, pass_result *pr)
{
    int rc;
    if (!force16 && open < 120 && match < 60 && -mismatch < 120) {
        rc = sw_striped_8(c1, o1, d1, m, c2, o2, d2, n, match, mismatch, open, ext, pr);
        if (rc) return rc;
        if (!pr->saturated) return 0;
    }
    rc = sw_striped_16(c1, o1, d1, m, c2, o2, d2, n, match, mismatch, open, ext, pr);
    if (rc) return rc;
    return pr->saturated ? 1 : 0;
}

int ora_striped_ssw_codes(const uint8_t *c1, int m, const uint8_t *c2, int n,
                          int match, int mismatch, int open, int ext, int force16, ora_ssw_result *res)
{
    pass_result f, r;
    if (m <= 0 || n <= 0) { memset(res, 0, sizeof(*res)); res->read_end1 = m - 1; return 0; }
    int rc = sw_striped_sat(c1, 0, 1, m, c2, 0, 1, n, match, mismatch, open, ext, force16, &f);
    if (rc) return rc;
    res->score1 = f.score; res->read_end1 = f.end_query; res->ref_end1 = f.end_ref;
    rc = sw_striped_sat(c1, f.end_query, -1, f.end_query + 1, c2, f.end_ref, -1, f.end_ref + 1,
                        match, mismatch, open, ext, force16, &r);
    if (rc) return rc;
    res->read_begin1 = f.end_query - r.end_query;
    res->ref_begin1 = f.end_ref - r.end_ref;
    return 0;
}

/* Batched, multi-threaded driver: a pthread pool pulling pairs from an atomic counter (dynamic schedule).
 * Returns the number of threads used, or a negative error code. */
typedef struct {
    const uint8_t *codes; const uint64_t *offsets; const uint32_t *lens;
    const uint32_t *tmpl_idx; const uint32_t *read_idx; uint64_t n_pairs;
    int match, mismatch, open, ext, force16;
    ora_ssw_result *out; int32_t *status;
    atomic_ullong next; atomic_int err;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *job = (batch_job *)arg;
    for (;;) {
        uint64_t p = atomic_fetch_add(&job->next, 1ULL);
        if (p >= job->n_pairs) break;
        uint32_t t = job->tmpl_idx[p], r = job->read_idx[p];
        int rc = ora_striped_ssw_codes(job->codes + job->offsets[t], (int)job->lens[t],
                                       job->codes + job->offsets[r], (int)job->lens[r],
                                       job->match, job->mismatch, job->open, job->ext, job->force16, &job->out[p]);
        if (rc < 0) atomic_store(&job->err, rc);
        if (job->status) job->status[p] = rc;
        if (rc == 1) {
            job->out[p].score1 = -1;
            job->out[p].ref_begin1 = job->out[p].ref_end1 = job->out[p].read_begin1 = job->out[p].read_end1 = -1;
        }
    }
    return NULL;
}

int ora_striped_ssw_batch(const uint8_t *codes, const uint64_t *offsets, const uint32_t *lens,
                          const uint32_t *tmpl_idx, const uint32_t *read_idx, uint64_t n_pairs,
                          int match, int mismatch, int open, int ext, int n_threads, int force16,
                          ora_ssw_result *out, int32_t *status)
{
    batch_job job;
    job.codes = codes; job.offsets = offsets; job.lens = lens; job.tmpl_idx = tmpl_idx; job.read_idx = read_idx;
    job.n_pairs = n_pairs; job.match = match; job.mismatch = mismatch; job.open = open; job.ext = ext;
    job.force16 = force16; job.out = out; job.status = status;
    atomic_init(&job.next, 0ULL); atomic_init(&job.err, 0);
    if (n_threads <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); n_threads = c > 0 ? (int)c : 1; }
    if ((uint64_t)n_threads > n_pairs) n_threads = n_pairs ? (int)n_pairs : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!th) return -1;
    int started = 0;
    for (int i = 1; i < n_threads; ++i)
        if (pthread_create(&th[started], NULL, batch_worker, &job) == 0) ++started;
    batch_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
    int err = atomic_load(&job.err);
    return err ? err : started + 1;
}

int ora_have_avx2(void)
{
    return __builtin_cpu_supports("avx2") ? 1 : 0;
}
