/*
 * nmsv_jump_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of the reference's breakpoint-refinement DP, nanomonsv/smith_waterman.py:5-127 (sw_jump), called from
 * nanomonsv/locate_bp.py:37,67. Unlike parasail this function IS in the reference checkout and is pure numpy, so the
 * restatement is pinned against vectors produced by the reference function itself (tests/golden/sw_jump_vectors.json,
 * generator tests/golden/make_golden_jump.py): PARITY PINNED for this row.
 *
 * Semantics kept exactly (they all change results):
 *   - H1 / H2 have zero boundaries and NO zero floor (values may go negative)          (:7-8, :16-23)
 *   - ties are resolved by tuple order: match, delete (i-1), insert (j-1), jump        (:21-23, :43-45)
 *   - the jump origin of row i is the FIRST maximum of H1_max[0..i-1], H1_max[r] = max over columns 0..m1 (column 0
 *     is the zero boundary), usable only if >= minimum_score                            (:25-26, :35-36)
 *   - jump = (H1_max + substitution) - jump_cost * ln(i - origin) in double precision; when it wins it is stored
 *     into the int32 matrix by C truncation                                            (:38-44)
 *   - end cell: last row if max(last row) >= max(last column), first maximum along it  (:54-59)
 *   - None unless a jump was taken, H1 part >= minimum_score and H2.max() - H1 part >= minimum_score   (:103-104)
 * The natural logarithms are passed in by the caller (numpy's log in the tests) so that no libm difference can leak.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t found;       /* 0 = the reference returns None */
    int32_t score;       /* H2.max() */
    int32_t i1_start, i1_end, i2_start, i2_end;   /* contig coordinates of the two fragments */
    int32_t j1_start, j1_end;                     /* region1 */
    int32_t j2_start, j2_end;                     /* region2 */
    uint32_t n_ops2, n_ops1;                      /* traceback operations (codes 1..4), H2 part then H1 part */
} ora_jump_result;

#define IDX(i, j, w) ((size_t)(i) * (size_t)(w) + (size_t)(j))

int ora_sw_jump(const uint8_t *contig, int n, const uint8_t *r1, int m1, const uint8_t *r2, int m2,
                int match, int mismatch_penalty, int gap_cost, int jump_cost, int minimum_score,
                const double *logtab /* logtab[d] = ln(d), d = 0..n */, ora_jump_result *res, uint8_t *ops)
{
    const int w1 = m1 + 1, w2 = m2 + 1;
    int32_t *H1 = (int32_t *)calloc((size_t)(n + 1) * w1, sizeof(int32_t));
    int32_t *H2 = (int32_t *)calloc((size_t)(n + 1) * w2, sizeof(int32_t));
    uint8_t *P1 = (uint8_t *)calloc((size_t)(n + 1) * w1, 1);
    uint8_t *P2 = (uint8_t *)calloc((size_t)(n + 1) * w2, 1);
    int32_t *hmax = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *hargmax = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    if (!H1 || !H2 || !P1 || !P2 || !hmax || !hargmax) { free(H1); free(H2); free(P1); free(P2); free(hmax); free(hargmax); return -1; }
    memset(res, 0, sizeof(*res));

    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= m1; ++j) {
            int mt = H1[IDX(i - 1, j - 1, w1)] + (contig[i - 1] == r1[j - 1] ? match : -mismatch_penalty);
            int de = H1[IDX(i - 1, j, w1)] - gap_cost;
            int in = H1[IDX(i, j - 1, w1)] - gap_cost;
            int best = mt, p = 1;
            if (de > best) { best = de; p = 2; }
            if (in > best) { best = in; p = 3; }
            H1[IDX(i, j, w1)] = best; P1[IDX(i, j, w1)] = (uint8_t)p;
        }
    for (int i = 0; i <= n; ++i) {
        int mx = H1[IDX(i, 0, w1)], am = 0;
        for (int j = 1; j <= m1; ++j) if (H1[IDX(i, j, w1)] > mx) { mx = H1[IDX(i, j, w1)]; am = j; }
        hmax[i] = mx; hargmax[i] = am;
    }
    int org = 0;   /* first argmax of hmax[0..i-1], maintained incrementally */
    for (int i = 1; i <= n; ++i) {
        if (i >= 2 && hmax[i - 1] > hmax[org]) org = i - 1;
        const int usable = hmax[org] >= minimum_score;
        for (int j = 1; j <= m2; ++j) {
            const int s = contig[i - 1] == r2[j - 1] ? match : -mismatch_penalty;
            int mt = H2[IDX(i - 1, j - 1, w2)] + s;
            int de = H2[IDX(i - 1, j, w2)] - gap_cost;
            int in = H2[IDX(i, j - 1, w2)] - gap_cost;
            int best = mt, p = 1;
            if (de > best) { best = de; p = 2; }
            if (in > best) { best = in; p = 3; }
            int32_t val = best;
            if (usable) {
                volatile double prod = (double)jump_cost * logtab[i - org];   /* no fused multiply-add */
                double jump = (double)(hmax[org] + s) - prod;
                if (jump > (double)best) { p = 4; val = (int32_t)jump; }
            }
            H2[IDX(i, j, w2)] = val; P2[IDX(i, j, w2)] = (uint8_t)p;
        }
    }
    int h2max = 0;
    for (size_t k = 0; k < (size_t)(n + 1) * w2; ++k) if (H2[k] > h2max) h2max = H2[k];

    int rowmax = H2[IDX(n, 0, w2)], rowarg = 0, colmax = H2[IDX(0, m2, w2)], colarg = 0;
    for (int j = 1; j <= m2; ++j) if (H2[IDX(n, j, w2)] > rowmax) { rowmax = H2[IDX(n, j, w2)]; rowarg = j; }
    for (int i = 1; i <= n; ++i) if (H2[IDX(i, m2, w2)] > colmax) { colmax = H2[IDX(i, m2, w2)]; colarg = i; }
    int ic, jc;
    if (rowmax >= colmax) { ic = n; jc = rowarg; } else { ic = colarg; jc = m2; }
    const int i2_end = ic, j2_end = jc;
    int jump_count = 0, h1_score = 0, i1_end = 0, j1_end = 0, i2_start = 0, j2_start = 0;
    uint32_t nops2 = 0, nops1 = 0;
    while (ic > 0 && jc > 0) {
        const int p = P2[IDX(ic, jc, w2)];
        if (ops) ops[nops2] = (uint8_t)p;
        ++nops2;
        if (p == 1) { --ic; --jc; }
        else if (p == 2) { --ic; }
        else if (p == 3) { --jc; }
        else if (p == 4) {
            /* origin of row ic */
            int o = 0;
            for (int r = 1; r < ic; ++r) if (hmax[r] > hmax[o]) o = r;
            i1_end = o; j1_end = hargmax[o]; h1_score = hmax[o]; i2_start = ic; j2_start = jc; jump_count = 1;
            break;
        } else break;
    }
    res->score = h2max; res->n_ops2 = nops2;
    if (h1_score < minimum_score || h2max - h1_score < minimum_score || jump_count == 0) {
        res->found = 0;
    } else {
        int i = i1_end, j = j1_end;
        while (i > 0 && j > 0) {
            const int p = P1[IDX(i, j, w1)];
            if (ops) ops[nops2 + nops1] = (uint8_t)p;
            ++nops1;
            if (p == 1) { --i; --j; }
            else if (p == 2) { --i; }
            else if (p == 3) { --j; }
            else break;
        }
        res->found = 1; res->n_ops1 = nops1;
        res->i1_start = i; res->i1_end = i1_end; res->i2_start = i2_start; res->i2_end = i2_end;
        res->j1_start = j + 1; res->j1_end = j1_end; res->j2_start = j2_start; res->j2_end = j2_end;
    }
    free(H1); free(H2); free(P1); free(P2); free(hmax); free(hargmax);
    return 0;
}
