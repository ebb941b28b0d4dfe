/*
 * nmsv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the arithmetic nanomonsv's supporting-read validation path delegates to
 * parasail (reference call sites: nanomonsv/count_sread_by_alignment.py:139,148-149 and
 * nanomonsv/long_read_validate.py:128,137-138):
 *
 *     M   = parasail.matrix_create("ACGT", 2, -2)
 *     res = parasail.ssw(template, read, 3, 1, M)   -> score1, ref_begin1, ref_end1, read_begin1, read_end1
 *
 * parasail itself (pip parasail==1.3.4, reference Dockerfile:49; C library v2.x) is NOT in /root/reference and
 * cannot be installed in this image, so this file restates its published algorithm (src/ssw.c,
 * src/sw_striped_*_{8,16}.c of upstream parasail):
 *   - Gotoh affine-gap local alignment, a gap of length k costs open + (k-1)*ext;
 *   - alphabet mapper case-insensitive, every byte outside the alphabet maps to a wildcard scoring 0;
 *   - end cell  = first column j (position in s2) whose maximum strictly raises the running maximum to its final
 *                 value, then the smallest row i (position in s1) in that column holding that value;
 *   - begin cell = the same rule applied to a second full local alignment of reverse(s1[0..end_i]) against
 *                 reverse(s2[0..end_j]);
 *   - NULL (here: return 1) when the 16-bit pass would saturate (any H > INT16_MAX - max(matrix) - 1).
 * PARITY UNPINNED against a real parasail binary (none is available); it is pinned against the hand-derived
 * golden vectors of tests/golden/  and cross-checked against an independent striped-SIMD
 * formulation (oracle/nmsv_striped.c) that follows parasail's own loop structure.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score1;
    int32_t ref_begin1;  /* begin in s2 (the long read), 0-based   */
    int32_t ref_end1;    /* end   in s2, 0-based inclusive         */
    int32_t read_begin1; /* begin in s1 (the template), 0-based    */
    int32_t read_end1;   /* end   in s1, 0-based inclusive         */
} ora_ssw_result;

/* parasail_matrix_create("ACGT", ...) mapper: case-insensitive, everything else -> wildcard index 4 */
uint8_t ora_base_code(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

void ora_encode(const char *s, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = ora_base_code((unsigned char)s[i]);
}

static inline int sub_score(uint8_t a, uint8_t b, int match, int mismatch)
{
    if (a > 3 || b > 3) return 0;
    return a == b ? match : mismatch;
}

#define NEG_INF (-(1 << 28))

/*
 * One local-alignment pass over code arrays c1 (rows, length m) and c2 (columns, length n).
 * The sequences are addressed as c1[o1 + d1*i], c2[o2 + d2*j] so the reversed second pass needs no copies.
 * Returns the best score and the end cell under the (first column, then smallest row) rule.
 * If the best score is 0 the end cell follows parasail's initialisation: end_ref = 0 column never set -> we
 * report (m-1, 0)-like garbage there too; callers must not interpret coordinates of a zero score.
 */
static void sw_pass(const uint8_t *c1, int64_t o1, int d1, int m,
                    const uint8_t *c2, int64_t o2, int d2, int n,
                    int match, int mismatch, int open, int ext,
                    int32_t *H, int32_t *E,
                    int *best_out, int *end_i_out, int *end_j_out)
{
    int best = 0, end_i = m - 1, end_j = 0;
    for (int i = 0; i < m; ++i) { H[i] = 0; E[i] = NEG_INF; }
    for (int j = 0; j < n; ++j) {
        uint8_t b = c2[o2 + (int64_t)d2 * j];
        int diag = 0;      /* H[i-1][j-1] */
        int up = 0;        /* H[i-1][j]   */
        int F = NEG_INF;   /* F[i-1][j]   */
        int colmax = 0, colmax_i = -1;
        for (int i = 0; i < m; ++i) {
            int hleft = H[i];
            int e = E[i] - ext; if (hleft - open > e) e = hleft - open;      /* E[i][j] */
            int f = F - ext;    if (up - open > f) f = up - open;            /* F[i][j] */
            int h = diag + sub_score(c1[o1 + (int64_t)d1 * i], b, match, mismatch);
            if (e > h) h = e;
            if (f > h) h = f;
            if (h < 0) h = 0;
            diag = hleft;
            H[i] = h; E[i] = e; F = f; up = h;
            if (h > colmax) { colmax = h; colmax_i = i; }   /* strict > keeps the smallest row */
        }
        if (colmax > best) { best = colmax; end_j = j; end_i = colmax_i; }
    }
    *best_out = best; *end_i_out = end_i; *end_j_out = end_j;
}

/*
 * parasail.ssw(s1, s2, open, ext, matrix_create("ACGT", match, mismatch)) on pre-encoded sequences.
 * Returns 0 and fills *res; returns 1 when parasail would return NULL (16-bit saturation); -1 on allocation failure.
 */
int ora_ssw_codes(const uint8_t *c1, int m, const uint8_t *c2, int n,
                  int match, int mismatch, int open, int ext, ora_ssw_result *res)
{
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1) * 2);
    if (!H) return -1;
    int32_t *E = H + (m > 0 ? m : 1);
    int best, ei, ej;
    sw_pass(c1, 0, 1, m, c2, 0, 1, n, match, mismatch, open, ext, H, E, &best, &ei, &ej);
    if (best > INT16_MAX - match - 1) { free(H); return 1; }
    res->score1 = best;
    res->read_end1 = ei;
    res->ref_end1 = ej;
    /* second pass on the reversed prefixes s1[0..ei], s2[0..ej] (parasail src/ssw.c "find the beginning") */
    int rbest, ri, rj;
    sw_pass(c1, ei, -1, ei + 1, c2, ej, -1, ej + 1, match, mismatch, open, ext, H, E, &rbest, &ri, &rj);
    res->read_begin1 = ei - ri;
    res->ref_begin1 = ej - rj;
    free(H);
    return 0;
}

int ora_ssw(const char *s1, int m, const char *s2, int n,
            int match, int mismatch, int open, int ext, ora_ssw_result *res)
{
    uint8_t *c = (uint8_t *)malloc((size_t)m + (size_t)n + 2);
    if (!c) return -1;
    ora_encode(s1, m, c);
    ora_encode(s2, n, c + m);
    int rc = ora_ssw_codes(c, m, c + m, n, match, mismatch, open, ext, res);
    free(c);
    return rc;
}

/* Batched form used by the tests and by the scalar CPU-baseline leg: pairs of sequences out of one code buffer. */
int ora_ssw_batch(const uint8_t *codes, const uint64_t *offsets, const uint32_t *lens,
                  const uint32_t *tmpl_idx, const uint32_t *read_idx, uint64_t n_pairs,
                  int match, int mismatch, int open, int ext, ora_ssw_result *out, int32_t *status)
{
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t t = tmpl_idx[p], r = read_idx[p];
        int rc = ora_ssw_codes(codes + offsets[t], (int)lens[t], codes + offsets[r], (int)lens[r],
                               match, mismatch, open, ext, &out[p]);
        if (rc < 0) return rc;
        if (status) status[p] = rc;
        if (rc == 1) { out[p].score1 = -1; out[p].ref_begin1 = out[p].ref_end1 = out[p].read_begin1 = out[p].read_end1 = -1; }
    }
    return 0;
}
