// nmsv_jump.cuh -- breakpoint-refinement DP of the reference: nanomonsv/smith_waterman.py:5-127 (sw_jump), called from
// nanomonsv/locate_bp.py:37,67. SURVEY.md section 8f "next" #2. Pinned against vectors produced by the reference's own
// function (tests/golden/sw_jump_vectors.json).
//
// Two linear-gap DP matrices without a zero floor over (contig x region1) and (contig x region2); the second one may
// "jump" from the best row of the first one at a cost of jump_cost * ln(distance) (double precision, truncated into
// the int32 matrix when it wins); then a traceback through H2 and H1. Semantics that change results are listed in
// oracle/nmsv_jump_oracle.c.
//
// Mapping: one thread block per problem. Cells of one anti-diagonal are independent: the block sweeps the diagonals,
// three rolling diagonal buffers in shared memory (indexed by the region position j), one byte of path per cell to a
// global workspace (read back only by the traceback), per-row "first maximum" of H1 through one packed 64-bit
// atomicMax per cell (value high, inverted column low: larger value wins, then smaller column). The jump origin of
// every row (first maximum of H1_max over the rows before it) is a block-wide prefix-maximum scan of the same packed
// words. This path is HBM/latency bound integer work (1 path byte per cell), not a DPX kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nmsv {

constexpr int kJumpThreads = 256;
constexpr int kJumpMaxRegion = 8190;      // 3 diagonal buffers of (m + 2) int32 must fit in shared memory

struct JumpProblem {
    uint64_t c_off, r1_off, r2_off;       // byte offsets of contig / region1 / region2 in `bytes`
    uint32_t n, m1, m2, res_idx;          // res_idx: position of this problem in the caller's arrays
    uint64_t p1_off, p2_off;              // path matrices in the byte workspace
    uint64_t row_off;                     // per-row arrays (index i = 0..n) in the 8-byte workspace
    uint64_t ops_off;                     // traceback operations of this problem in `ops`
};

struct JumpResult {                      // mirrors nmsv_jump_result (include/nmsv.h)
    int32_t found, score;
    int32_t i1_start, i1_end, i2_start, i2_end;
    int32_t j1_start, j1_end, j2_start, j2_end;
    uint32_t n_ops2, n_ops1;
};

struct JumpParams {
    const uint8_t* bytes;
    const JumpProblem* problems;
    uint32_t n_problems;
    uint8_t* paths;                       // byte workspace
    unsigned long long* rows;             // 8-byte workspace: per problem 4 arrays of (n + 1): rowbest, prefix, prod, aux
    const double* logtab;
    uint32_t logtab_len;
    JumpResult* results;
    uint8_t* ops;
    int match, mismatch_penalty, gap_cost, jump_cost, minimum_score;
};

__device__ __forceinline__ unsigned long long jpack(int value, uint32_t index)
{
    return ((unsigned long long)(uint32_t)(value + (1 << 30)) << 32) | (unsigned long long)(0xffffffffu - index);
}
__device__ __forceinline__ int junpack_value(unsigned long long w) { return (int)(uint32_t)(w >> 32) - (1 << 30); }
__device__ __forceinline__ uint32_t junpack_index(unsigned long long w) { return 0xffffffffu - (uint32_t)w; }

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* red)
{
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = red[0];
    for (int w = 1; w < kJumpThreads / 32; ++w) r = red[w] > r ? red[w] : r;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kJumpThreads) sw_jump_kernel(const JumpParams p)
{
    extern __shared__ __align__(16) int32_t jsm[];
    __shared__ unsigned long long red[kJumpThreads / 32];
    __shared__ unsigned long long chunk_max[kJumpThreads];
    __shared__ int h2max_s;
    const int tid = threadIdx.x;

    for (uint32_t pi = blockIdx.x; pi < p.n_problems; pi += gridDim.x) {
        const JumpProblem pr = p.problems[pi];
        const int n = (int)pr.n, m1 = (int)pr.m1, m2 = (int)pr.m2;
        const uint8_t* contig = p.bytes + pr.c_off;
        const uint8_t* reg1 = p.bytes + pr.r1_off;
        const uint8_t* reg2 = p.bytes + pr.r2_off;
        uint8_t* P1 = p.paths + pr.p1_off;
        uint8_t* P2 = p.paths + pr.p2_off;
        unsigned long long* rowbest = p.rows + pr.row_off;              // packed (H1_max[i], first argmax column)
        unsigned long long* prefix = rowbest + (n + 1);                 // packed first maximum of rowbest over rows <= i
        double* prod = reinterpret_cast<double*>(prefix + (n + 1));     // jump_cost * ln(i - origin(i))
        int32_t* lastcol = reinterpret_cast<int32_t*>(prod + (n + 1));  // H2[i][m2]
        int32_t* lastrow = lastcol + (n + 1);                           // H2[n][j], j = 0..m2 (m2 + 1 <= n + 1 not assumed:
                                                                        // the host reserves max(n, m2) + 1 entries)
        const int mbuf = (m1 > m2 ? m1 : m2) + 2;
        int32_t* d0 = jsm;
        int32_t* d1 = jsm + mbuf;
        int32_t* d2 = jsm + 2 * mbuf;

        // ---------------------------------------------------------------- H1
        for (int i = tid; i <= n; i += kJumpThreads) rowbest[i] = jpack(0, 0);      // column 0 is the zero boundary
        for (int j = tid; j < mbuf; j += kJumpThreads) { d0[j] = 0; d1[j] = 0; d2[j] = 0; }
        __syncthreads();
        {
            int32_t *cur = d0, *prev1 = d1, *prev2 = d2;     // diagonals d, d-1, d-2 ; entry j <-> cell (d - j, j)
            for (int d = 2; d <= n + m1; ++d) {
                const int jlo = d - n > 1 ? d - n : 1;
                const int jhi = d - 1 < m1 ? d - 1 : m1;
                for (int j = jlo + tid; j <= jhi; j += kJumpThreads) {
                    const int i = d - j;
                    const int s = contig[i - 1] == reg1[j - 1] ? p.match : -p.mismatch_penalty;
                    const int mt = prev2[j - 1] + s;
                    const int de = prev1[j] - p.gap_cost;         // cell (i-1, j)
                    const int in = prev1[j - 1] - p.gap_cost;     // cell (i, j-1)
                    int best = mt, path = 1;
                    if (de > best) { best = de; path = 2; }
                    if (in > best) { best = in; path = 3; }
                    cur[j] = best;
                    P1[(size_t)i * (m1 + 1) + j] = (uint8_t)path;
                    atomicMax(&rowbest[i], jpack(best, (uint32_t)j));
                }
                if (tid == 0) {
                    if (d <= n) cur[0] = 0;        // cell (d, 0)
                    if (d <= m1) cur[d] = 0;       // cell (0, d)
                }
                __syncthreads();
                int32_t* t = prev2; prev2 = prev1; prev1 = cur; cur = t;
            }
        }
        __threadfence_block();
        __syncthreads();

        // ---------------------------------------------------------------- jump origins: prefix first-maximum over rows
        {
            const int per = (n + 1 + kJumpThreads - 1) / kJumpThreads;
            const int lo = tid * per, hi = (lo + per < n + 1) ? lo + per : n + 1;
            unsigned long long run = 0;
            // a row with a larger H1_max wins; equal maxima keep the EARLIER row: repack as (value, inverted row)
            for (int i = lo; i < hi; ++i) {
                const unsigned long long w = jpack(junpack_value(rowbest[i]), (uint32_t)i);
                run = w > run ? w : run;
            }
            chunk_max[tid] = run;
            __syncthreads();
            unsigned long long before = 0;
            for (int q = 0; q < tid; ++q) before = chunk_max[q] > before ? chunk_max[q] : before;
            run = before;
            for (int i = lo; i < hi; ++i) {
                const unsigned long long w = jpack(junpack_value(rowbest[i]), (uint32_t)i);
                run = w > run ? w : run;
                prefix[i] = run;
            }
            __syncthreads();
            for (int i = 1 + tid; i <= n; i += kJumpThreads) {
                const uint32_t org = junpack_index(prefix[i - 1]);
                const uint32_t dist = (uint32_t)i - org;
                const double lg = dist < p.logtab_len ? p.logtab[dist] : log((double)dist);
                prod[i] = __dmul_rn((double)p.jump_cost, lg);
            }
            if (tid == 0) h2max_s = 0;
            __syncthreads();
        }

        // ---------------------------------------------------------------- H2
        for (int j = tid; j < mbuf; j += kJumpThreads) { d0[j] = 0; d1[j] = 0; d2[j] = 0; }
        for (int i = tid; i <= n; i += kJumpThreads) lastcol[i] = 0;
        for (int j = tid; j <= m2; j += kJumpThreads) lastrow[j] = 0;
        __syncthreads();
        {
            int32_t *cur = d0, *prev1 = d1, *prev2 = d2;
            int local_max = 0;
            for (int d = 2; d <= n + m2; ++d) {
                const int jlo = d - n > 1 ? d - n : 1;
                const int jhi = d - 1 < m2 ? d - 1 : m2;
                for (int j = jlo + tid; j <= jhi; j += kJumpThreads) {
                    const int i = d - j;
                    const int s = contig[i - 1] == reg2[j - 1] ? p.match : -p.mismatch_penalty;
                    const int mt = prev2[j - 1] + s;
                    const int de = prev1[j] - p.gap_cost;
                    const int in = prev1[j - 1] - p.gap_cost;
                    int best = mt, path = 1;
                    if (de > best) { best = de; path = 2; }
                    if (in > best) { best = in; path = 3; }
                    const int hm = junpack_value(prefix[i - 1]);              // H1_max at the origin row
                    if (hm >= p.minimum_score) {
                        const double jump = __dsub_rn((double)(hm + s), prod[i]);
                        if (jump > (double)best) { path = 4; best = (int)jump; }   // C truncation, like the int32 store
                    }
                    cur[j] = best;
                    P2[(size_t)i * (m2 + 1) + j] = (uint8_t)path;
                    local_max = best > local_max ? best : local_max;
                    if (i == n) lastrow[j] = best;
                    if (j == m2) lastcol[i] = best;
                }
                if (tid == 0) {
                    if (d <= n) cur[0] = 0;
                    if (d <= m2) cur[d] = 0;
                }
                __syncthreads();
                int32_t* t = prev2; prev2 = prev1; prev1 = cur; cur = t;
            }
            atomicMax(&h2max_s, local_max);
        }
        __threadfence_block();
        __syncthreads();

        // ---------------------------------------------------------------- end cell: first maxima of last row / column
        unsigned long long rw = 0, cw = 0;
        for (int j = tid; j <= m2; j += kJumpThreads) { const unsigned long long w = jpack(lastrow[j], (uint32_t)j); rw = w > rw ? w : rw; }
        for (int i = tid; i <= n; i += kJumpThreads) { const unsigned long long w = jpack(lastcol[i], (uint32_t)i); cw = w > cw ? w : cw; }
        rw = block_max_u64(rw, red);
        cw = block_max_u64(cw, red);

        // ---------------------------------------------------------------- traceback (one thread)
        if (tid == 0) {
            JumpResult res;
            res.found = 0; res.score = h2max_s;
            res.i1_start = res.i1_end = res.i2_start = res.i2_end = 0;
            res.j1_start = res.j1_end = res.j2_start = res.j2_end = 0;
            res.n_ops2 = res.n_ops1 = 0;
            int ic, jc;
            if (junpack_value(rw) >= junpack_value(cw)) { ic = n; jc = (int)junpack_index(rw); }
            else { ic = (int)junpack_index(cw); jc = m2; }
            const int i2_end = ic, j2_end = jc;
            uint8_t* ops = p.ops + pr.ops_off;
            uint32_t n2 = 0, n1 = 0;
            int jumped = 0, h1_score = 0, i1_end = 0, j1_end = 0, i2_start = 0, j2_start = 0;
            while (ic > 0 && jc > 0) {
                const int path = P2[(size_t)ic * (m2 + 1) + jc];
                ops[n2++] = (uint8_t)path;
                if (path == 1) { --ic; --jc; }
                else if (path == 2) { --ic; }
                else if (path == 3) { --jc; }
                else if (path == 4) {
                    const unsigned long long w = prefix[ic - 1];
                    i1_end = (int)junpack_index(w);
                    h1_score = junpack_value(w);
                    j1_end = (int)junpack_index(rowbest[i1_end]);
                    i2_start = ic; j2_start = jc; jumped = 1;
                    break;
                } else break;
            }
            res.n_ops2 = n2;
            if (jumped && h1_score >= p.minimum_score && res.score - h1_score >= p.minimum_score) {
                int i = i1_end, j = j1_end;
                while (i > 0 && j > 0) {
                    const int path = P1[(size_t)i * (m1 + 1) + j];
                    ops[n2 + n1] = (uint8_t)path;
                    ++n1;
                    if (path == 1) { --i; --j; }
                    else if (path == 2) { --i; }
                    else if (path == 3) { --j; }
                    else break;
                }
                res.found = 1; res.n_ops1 = n1;
                res.i1_start = i; res.i1_end = i1_end; res.i2_start = i2_start; res.i2_end = i2_end;
                res.j1_start = j + 1; res.j1_end = j1_end; res.j2_start = j2_start; res.j2_end = j2_end;
            }
            p.results[pr.res_idx] = res;
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Second cut: one WARP per problem (regions up to kJumpWarpMaxRegion columns; longer ones keep the block kernel above).
//
// Lane l owns the strip of sl = ceil(m / 32) consecutive region columns and keeps the previous row of its strip in
// registers; at step t it computes row t - l (anti-diagonal wavefront over the lanes). One shuffle per step hands the
// strip's last cell to the next lane (its "left" neighbour now, its diagonal one step later), a second one carries the
// running (row maximum, first column) of H1 along the row, so H1_max / H1_argmax fall out of the last strip without
// atomics. The path of a strip row is one packed 64-bit word (2 bits per cell): 256 bytes per row and warp, coalesced,
// instead of one byte per cell. The jump term is two values per row (the cell matches or not): they are precomputed
// per row as (ceil, trunc) integers of the reference's double expression -- x > b <=> ceil(x) > b for an integer b --
// so the inner loop stays in integers. Prefix maxima over rows by a warp scan; traceback by lane 0 over the packed words.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kJumpWarpMaxRegion = 1024;
constexpr int kJumpWarpsPerBlock = 4;

struct JumpWarpProblem {
    uint64_t c_off, r1_off, r2_off;
    uint32_t n, m1, m2, res_idx;
    uint64_t w1_off, w2_off;              // packed path words, (n + 1) x 32 each, in the 8-byte workspace
    uint64_t row_off;                     // per-row arrays in the 8-byte workspace: rowbest, prefix (n + 1 each), rowinfo (2 x (n + 1))
    uint64_t ops_off;
};

struct JumpWarpParams {
    const uint8_t* bytes;
    const JumpWarpProblem* problems;
    uint32_t n_problems;
    uint32_t* next;                       // work counter (zeroed before the launch)
    unsigned long long* work;             // 8-byte workspace
    const double* logtab;
    uint32_t logtab_len;
    JumpResult* results;
    uint8_t* ops;
    int match, mismatch_penalty, gap_cost, jump_cost, minimum_score;
};

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    return v;
}

// one DP matrix of a problem. kJump: the second matrix (jump term, end-cell bookkeeping).
template <int SMAX, bool kJump>
__device__ __forceinline__ void jump_matrix(const JumpWarpParams& p, const uint8_t* __restrict__ contig, const uint8_t* __restrict__ reg,
                                            int n, int m, unsigned long long* __restrict__ words,
                                            unsigned long long* __restrict__ rowbest, const int4* __restrict__ rowinfo,
                                            int& h2max, unsigned long long& lastrow, unsigned long long& lastcol)
{
    const int lane = threadIdx.x & 31;
    const int sl = m > 0 ? (m + 31) / 32 : 1;
    const int first = 1 + lane * sl;                       // first column of this lane's strip (1-based)
    const int lane_last = m > 0 ? (m - 1) / sl : 0;        // lane that owns column m (its last valid cell)
    int nvalid = m - first + 1;                            // cells of this lane's strip inside the matrix
    nvalid = nvalid < 0 ? 0 : (nvalid > sl ? sl : nvalid);
    uint32_t rc[SMAX / 4];                                 // region bases of the strip, four per register
    int hprev[SMAX];
#pragma unroll
    for (int q = 0; q < SMAX / 4; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int k = 4 * q + b;
            w |= (uint32_t)(k < nvalid ? reg[first + k - 1] : 0) << (8 * b);
        }
        rc[q] = w;
    }
#pragma unroll
    for (int k = 0; k < SMAX; ++k) hprev[k] = 0;
    int diag_keep = 0, last_out = 0;
    int rbv_out = 0, rbj_out = 0;                          // running (row maximum, its first column) handed along the row
    int local_max = 0;
    unsigned long long lrow = jpack(0, 0), lcol = jpack(0, 0);
    const int g = p.gap_cost, ms = p.match, mm = -p.mismatch_penalty;
    // the contig base and the jump values of the NEXT row are loaded one step ahead (L2 latency off the dependent chain)
    uint32_t c_next = 0;
    int4 info_next = make_int4(0, 0, 0, 0);
    {
        const int i0 = 1 - lane;
        if (i0 >= 1 && i0 <= n) { c_next = contig[i0 - 1]; if (kJump) info_next = __ldcg(rowinfo + i0); }
    }
    for (int t = 0; t < n + 31; ++t) {
        int left_in = __shfl_up_sync(0xffffffffu, last_out, 1);
        int rbv = __shfl_up_sync(0xffffffffu, rbv_out, 1);
        int rbj = __shfl_up_sync(0xffffffffu, rbj_out, 1);
        if (lane == 0) { left_in = 0; rbv = 0; rbj = 0; }       // column 0 of every row holds 0
        const int i = t - lane + 1;
        const uint32_t c = c_next;
        const int4 info = info_next;
        if (i + 1 >= 1 && i + 1 <= n) {
            c_next = contig[i];
            if (kJump) info_next = __ldcg(rowinfo + i + 1);   // written earlier in this kernel: not the read-only path
        }
        if (i >= 1 && i <= n) {
            int diag = diag_keep, left = left_in;
            unsigned long long word = 0;
#pragma unroll
            for (int k = 0; k < SMAX; ++k) {
                if (k < nvalid) {
                    const bool eq = c == ((rc[k >> 2] >> (8 * (k & 3))) & 0xffu);
                    const int mt = diag + (eq ? ms : mm);
                    const int de = hprev[k] - g;
                    const int in = left - g;
                    // The only value that depends on the previous cell of the row is `in`: the dependent chain per cell is one
                    // add and one or two max instructions; the path (first of match, delete, insert that holds the maximum;
                    // then the jump if it is strictly larger) is worked out beside the chain.
                    int best = __vimax3_s32(mt, de, in);
                    int path = mt == best ? 0 : (de == best ? 1 : 2);      // stored as path - 1
                    if (kJump) {
                        // jump > best <=> ceil(jump) > best for an integer best; and whenever it is, trunc(jump) >= best,
                        // otherwise trunc(jump) <= best: the new value is max(best, trunc(jump)) in every case
                        path = (eq ? info.x : info.z) > best ? 3 : path;
                        best = max(best, eq ? info.y : info.w);
                        local_max = max(local_max, best);
                    } else if (best > rbv) { rbv = best; rbj = first + k; }      // strict: the first column keeps a tie
                    diag = hprev[k];
                    hprev[k] = best;
                    left = best;
                    word |= (unsigned long long)path << (2 * k);
                }
            }
            words[(size_t)i * 32 + lane] = word;
            if (!kJump && lane == lane_last) rowbest[i] = jpack(rbv, (uint32_t)rbj);
            if (kJump) {
                if (i == n) {                                 // last row: first maximum over this strip
#pragma unroll
                    for (int k = 0; k < SMAX; ++k)
                        if (k < nvalid) { const unsigned long long w = jpack(hprev[k], (uint32_t)(first + k)); lrow = w > lrow ? w : lrow; }
                }
                if (lane == lane_last && nvalid > 0) {        // last column: the strip's last valid cell is column m
                    const unsigned long long w = jpack(left, (uint32_t)i);
                    lcol = w > lcol ? w : lcol;
                }
            }
            last_out = left;
            rbv_out = rbv; rbj_out = rbj;
            diag_keep = left_in;
        }
    }
    if (kJump) {
        local_max = (int)(junpack_value(warp_max_u64(jpack(local_max, 0))));
        h2max = local_max;
        lastrow = warp_max_u64(lrow);
        lastcol = __shfl_sync(0xffffffffu, lcol, lane_last);
    }
}

template <int SMAX>
__device__ __forceinline__ void jump_warp_problem(const JumpWarpParams& p, const JumpWarpProblem& pr)
{
    const int lane = threadIdx.x & 31;
    const int n = (int)pr.n, m1 = (int)pr.m1, m2 = (int)pr.m2;
    const uint8_t* contig = p.bytes + pr.c_off;
    unsigned long long* W1 = p.work + pr.w1_off;
    unsigned long long* W2 = p.work + pr.w2_off;
    unsigned long long* rowbest = p.work + pr.row_off;
    unsigned long long* prefix = rowbest + (n + 1);
    int4* rowinfo = reinterpret_cast<int4*>(prefix + (n + 1));      // 2 (n + 1) words after an even offset: 16-byte aligned
    int h2max = 0;
    unsigned long long rw = 0, cw = 0;

    if (lane == 0) rowbest[0] = jpack(0, 0);
    if (m1 == 0) for (int i = 1 + lane; i <= n; i += 32) rowbest[i] = jpack(0, 0);
    jump_matrix<SMAX, false>(p, contig, p.bytes + pr.r1_off, n, m1, W1, rowbest, nullptr, h2max, rw, cw);
    __threadfence();
    __syncwarp();
    // ---- jump origins: first maximum of H1_max over the rows before i (a row with a larger maximum wins, equal maxima keep
    //      the EARLIER row), then the two jump values of every row
    {
        unsigned long long carry = 0;
        for (int base = 0; base <= n; base += 32) {
            const int i = base + lane;
            unsigned long long w = i <= n ? jpack(junpack_value(__ldcg(rowbest + i)), (uint32_t)i) : 0ull;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w = o > w ? o : w;
            }
            w = carry > w ? carry : w;
            if (i <= n) prefix[i] = w;
            carry = __shfl_sync(0xffffffffu, w, 31);
        }
        __threadfence();
        __syncwarp();
        for (int i = 1 + lane; i <= n; i += 32) {
            const unsigned long long w = __ldcg(prefix + i - 1);
            const int hm = junpack_value(w);
            int4 info = make_int4(-0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1);      // no jump from this row
            if (hm >= p.minimum_score) {
                const uint32_t dist = (uint32_t)i - junpack_index(w);
                const double lg = dist < p.logtab_len ? p.logtab[dist] : log((double)dist);
                const double prod = __dmul_rn((double)p.jump_cost, lg);
                const double jm = __dsub_rn((double)(hm + p.match), prod), jx = __dsub_rn((double)(hm - p.mismatch_penalty), prod);
                info = make_int4((int)ceil(jm), (int)jm, (int)ceil(jx), (int)jx);
            }
            rowinfo[i] = info;
        }
        __threadfence();
        __syncwarp();
    }
    jump_matrix<SMAX, true>(p, contig, p.bytes + pr.r2_off, n, m2, W2, rowbest, rowinfo, h2max, rw, cw);
    __threadfence();
    __syncwarp();

    // ---- traceback (lane 0) over the packed path words
    if (lane == 0) {
        JumpResult res;
        res.found = 0; res.score = h2max;
        res.i1_start = res.i1_end = res.i2_start = res.i2_end = 0;
        res.j1_start = res.j1_end = res.j2_start = res.j2_end = 0;
        res.n_ops2 = res.n_ops1 = 0;
        const int sl1 = m1 > 0 ? (m1 + 31) / 32 : 1, sl2 = m2 > 0 ? (m2 + 31) / 32 : 1;
        auto path_at = [](const unsigned long long* W, int sl, int i, int j) {
            const int l = (j - 1) / sl, k = (j - 1) - l * sl;
            return (int)((__ldcg(W + (size_t)i * 32 + l) >> (2 * k)) & 3ull) + 1;
        };
        int ic, jc;
        if (junpack_value(rw) >= junpack_value(cw)) { ic = n; jc = (int)junpack_index(rw); }
        else { ic = (int)junpack_index(cw); jc = m2; }
        const int i2_end = ic, j2_end = jc;
        uint8_t* ops = p.ops + pr.ops_off;
        uint32_t n2 = 0, n1 = 0;
        int jumped = 0, h1_score = 0, i1_end = 0, j1_end = 0, i2_start = 0, j2_start = 0;
        while (ic > 0 && jc > 0) {
            const int path = path_at(W2, sl2, ic, jc);
            ops[n2++] = (uint8_t)path;
            if (path == 1) { --ic; --jc; }
            else if (path == 2) { --ic; }
            else if (path == 3) { --jc; }
            else {
                const unsigned long long w = __ldcg(prefix + ic - 1);
                i1_end = (int)junpack_index(w);
                h1_score = junpack_value(w);
                j1_end = (int)junpack_index(__ldcg(rowbest + i1_end));
                i2_start = ic; j2_start = jc; jumped = 1;
                break;
            }
        }
        res.n_ops2 = n2;
        if (jumped && h1_score >= p.minimum_score && res.score - h1_score >= p.minimum_score) {
            int i = i1_end, j = j1_end;
            while (i > 0 && j > 0) {
                const int path = path_at(W1, sl1, i, j);
                ops[n2 + n1] = (uint8_t)path;
                ++n1;
                if (path == 1) { --i; --j; }
                else if (path == 2) { --i; }
                else { --j; }
            }
            res.found = 1; res.n_ops1 = n1;
            res.i1_start = i; res.i1_end = i1_end; res.i2_start = i2_start; res.i2_end = i2_end;
            res.j1_start = j + 1; res.j1_end = j1_end; res.j2_start = j2_start; res.j2_end = j2_end;
        }
        p.results[pr.res_idx] = res;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kJumpWarpsPerBlock * 32, 4) sw_jump_warp_kernel(const JumpWarpParams p)
{
    for (;;) {
        uint32_t pi = 0;
        if ((threadIdx.x & 31) == 0) pi = atomicAdd(p.next, 1u);
        pi = __shfl_sync(0xffffffffu, pi, 0);
        if (pi >= p.n_problems) break;
        const JumpWarpProblem pr = p.problems[pi];
        const int mx = (int)(pr.m1 > pr.m2 ? pr.m1 : pr.m2);
        if (mx <= 256) jump_warp_problem<8>(p, pr);
        else if (mx <= 512) jump_warp_problem<16>(p, pr);
        else jump_warp_problem<32>(p, pr);
    }
}

}  // namespace nmsv
