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
// Warp kernel: one WARP per problem (regions up to kJumpWarpMaxRegion columns; longer ones keep the block kernel above).
//
// Lane l owns a strip of consecutive region columns; at step t it computes row t - l + 1 (anti-diagonal wavefront over
// the lanes). One shuffle per step hands the strip's last cell to the next lane (its "left" neighbour now, its diagonal
// one step later), two more carry the running (row maximum, first column) of H1 along the row, so H1_max / H1_argmax
// fall out of lane 31 without atomics.
//
// Third cut (round 2): straight-line cell code, ONE small body for every region length.
//   * Strips: S = ceil(m / 32); lanes 0..a-1 own S columns, lanes a..31 own S-1 (a = m - 32(S-1)): every lane is busy.
//     The previous row of a strip lives in shared memory as groups of four cells ([group][lane] uint4, conflict free);
//     the step walks the ceil(S/4) groups of its strip: one LDS.128 / STS.128 and one word of four region bases per group,
//     the four cells of a group fully unrolled in registers. Cells beyond a lane's own columns (at most 4, all in the
//     last group) are phantoms: they are computed like any other cell and dropped by selects in the last group only.
//     (Strips held entirely in registers need one instantiation per S; with 32 of them -- or two steps per loop
//     iteration -- the warps of an SM run a dozen different loop bodies side by side and the kernel becomes bound by
//     instruction fetch: measured 34 ms for the 3000-problem set against 3.6 ms when every problem has the same S.)
//   * All values are kept times 32 and minus the gap cost (G = 32*H - 32*g): "delete" and "insert" then need no
//     instruction (G of the cell above / to the left is already H - g), and the first row maximum of a group is ONE DPX
//     instruction per cell, max(best + (3 - c), run): the low bits carry the cell index, the first cell wins a tie.
//     Per cell of H1: equality predicate, select of the substitution score, 1 add, 2 max (dependent chain: max, add),
//     the re-bias add, 2 x (subtract + funnel shift) for the two path bits, 1 DPX = 11 instructions; H2 adds the jump
//     (max3 instead of max, 2 selects, 1 max, subtract + funnel shift) = 16. The previous cut spent 28 with a divergent
//     branch per cell.
//   * Path bits are sign bits shifted into per-row plane words (del > match, insert > both, jump > all), combined once per
//     step into two planes (path - 1 = low + 2 * high) and stored as one 8-byte word per lane and STEP: the 32 lanes of a
//     step write one contiguous 256-byte line.
//   * The jump term is two values per row (the cell matches or not), precomputed per row in double precision with the
//     caller's numpy.log table like the reference, as integers (ceil, trunc) times 32: jump > best <=> ceil(jump) > best
//     for an integer best, and the stored value is max(best, trunc(jump)) in every case.
//   * Traceback by the whole warp in lock step: the lanes fetch the plane words of 32 consecutive rows of the current
//     strip at once, the walk then reads them by shuffle (one L2 round trip per <= 32 operations instead of one each).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kJumpWarpMaxRegion = 1024;
constexpr int kJumpWarpsPerBlock = 4;
constexpr int kJumpScaleShift = 5;                 // values times 32: five low bits for the cell index of a strip
constexpr int kJumpNever = -(1 << 30);             // below every scaled value the host admits (|32 H| < 2^29)
constexpr int kJumpGroups = kJumpWarpMaxRegion / 32 / 4;      // groups of four cells per strip
constexpr size_t kJumpWarpSmem = (size_t)kJumpGroups * 32 * (16 + 4);       // previous row + region bases of one warp

struct JumpWarpProblem {
    uint64_t c_off, r1_off, r2_off;
    uint32_t n, m1, m2, res_idx;
    uint64_t w1_off, w2_off;              // path plane words, (n + 32) x 32 each, in the 8-byte workspace (indexed by step, lane)
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
    uint8_t* ops;                         // traceback operations, per problem at its (upper-bound) ops_off
    int match, mismatch_penalty, gap_cost, jump_cost, minimum_score;
};

__device__ __noinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    return v;
}

// (w << 1) | sign(d): collects "d < 0" bits, the first cell ends up in the highest of the bits
__device__ __forceinline__ uint32_t push_sign(uint32_t w, int d) { return __funnelshift_l((uint32_t)d, w, 1); }

struct JumpMatrixOut { int h2max; unsigned long long lastrow, lastcol; };

// one DP matrix of a problem, m >= 1 columns. kJump: the second matrix (jump term, end-cell bookkeeping).
// gs / rcs: this warp's shared memory ([group][lane] uint4 previous row, [group][lane] four region bases).
template <bool kJump>
__device__ __noinline__ JumpMatrixOut jump_matrix(const JumpWarpParams& p, const uint8_t* __restrict__ contig,
                                                  const uint8_t* __restrict__ reg, const int n, const int m, uint2* __restrict__ words,
                                                  unsigned long long* __restrict__ rowbest, const int4* __restrict__ rowinfo,
                                                  int4* __restrict__ gs, uint32_t* __restrict__ rcs)
{
    constexpr int kScale = 1 << kJumpScaleShift;
    const int lane = threadIdx.x & 31;
    const int S = (m + 31) / 32, ng = (S + 3) / 4;
    const int a = m - 32 * (S - 1);                        // lanes 0..a-1 own S columns, the others S-1 (1 <= a <= 32)
    const bool shortl = lane >= a;
    const int first = 1 + lane * S - (shortl ? lane - a : 0);      // first column of this lane's strip (1-based)
    const int nreal = shortl ? S - 1 : S;
    const int kbase = 4 * (ng - 1);                        // first cell of the last group; kbase <= nreal <= kbase + 4
    const int dcap = nreal - kbase;                        // cells of the last group that are this lane's own
    const int g32 = p.gap_cost * kScale;
    const int sge = p.match * kScale + g32, sgn = g32 - p.mismatch_penalty * kScale;
    int4* const gl = gs + lane;
    uint32_t* const rl = rcs + lane;
    for (int q = 0; q < ng; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int k = 4 * q + b;
            if (k < nreal) w |= (uint32_t)reg[first + k - 1] << (8 * b);
        }
        rl[q * 32] = w;
        gl[q * 32] = make_int4(-g32, -g32, -g32, -g32);    // G = 32 H - 32 g of the previous row
    }
    __syncwarp();
    int diag_keep = -g32, last_out = -g32;
    int rbv_out = 0, rbj_out = 0;                          // running (row maximum x 32, its first column) handed along the row
    int lmax = 0, lc_v = 0, lc_i = 0;
    // the contig base and the jump values of the NEXT row are loaded one step ahead (L2 latency off the dependent chain)
    uint32_t c_next = 0;
    int4 info_next = make_int4(kJumpNever, kJumpNever, kJumpNever, kJumpNever);
    {
        const int i0 = 1 - lane;
        if (i0 >= 1 && i0 <= n) { c_next = contig[i0 - 1]; if (kJump) info_next = __ldcg(rowinfo + i0); }
    }
    // one cell: diag / left / planes / running values are updated in place; returns the new G of the cell.
    // hprev = 32 H of the cell to the left (left = hprev - g32): the only dependent chain along a row is ONE DPX
    // instruction per cell, h = max(hprev - g, bb); everything else hangs off it.
    int diag, left, hprev, run, lm;
    uint32_t wdel, wins, wj;
    int4 info = make_int4(kJumpNever, kJumpNever, kJumpNever, kJumpNever);
    const int mg32 = -g32;
    auto cell = [&](const uint32_t x, const int b, const int de, const int kk) -> int {
        const bool eq = ((x >> (8 * b)) & 0xffu) == 0u;
        const int mt = diag + (eq ? sge : sgn);              // H(i-1, j-1) + s
        const int bb = max(mt, de);                          // de = H(i-1, j) - g
        int best;
        if (kJump) {
            const int bj = max(bb, eq ? info.y : info.w);    // trunc(jump) joins before the chain
            best = __viaddmax_s32(hprev, mg32, bj);
            wj = push_sign(wj, max(bb, left) - (eq ? info.x : info.z));      // jump iff ceil(jump) > max(match, delete, insert)
        } else {
            best = __viaddmax_s32(hprev, mg32, bb);
            run = __viaddmax_s32(best, kk, run);             // first maximum of the strip: cell index in the low bits
        }
        wdel = push_sign(wdel, mt - de);                     // delete only if strictly larger than match
        wins = push_sign(wins, bb - left);                   // insert only if strictly larger than both
        diag = de;
        hprev = best;
        left = best - g32;
        return left;
    };
#pragma unroll 1
    for (int t = 0; t < n + 31; ++t) {
        int left_in = __shfl_up_sync(0xffffffffu, last_out, 1);
        int rbv = 0, rbj = 0;
        if (!kJump) { rbv = __shfl_up_sync(0xffffffffu, rbv_out, 1); rbj = __shfl_up_sync(0xffffffffu, rbj_out, 1); }
        if (lane == 0) { left_in = -g32; rbv = 0; rbj = 0; }       // column 0 of every row holds 0
        const int i = t - lane + 1;
        const uint32_t c4 = c_next * 0x01010101u;
        info = info_next;
        if (i + 1 >= 1 && i + 1 <= n) {
            c_next = contig[i];
            if (kJump) info_next = __ldcg(rowinfo + i + 1);        // written earlier in this kernel: not the read-only path
        }
        if (i >= 1 && i <= n) {
            diag = diag_keep; left = left_in; hprev = left_in + g32; run = kJumpNever; lm = kJumpNever;
            wdel = 0; wins = 0; wj = 0;
            int4 g = gl[0];
            uint32_t rcw = rl[0];
            int kq = 28;                                     // 31 - k of the group's first cell, minus 3
#pragma unroll 2
            for (int q = 0; q < ng - 1; ++q) {               // whole groups: every cell is this lane's own
                const int4 gn = gl[(q + 1) * 32];            // next group, off the dependent chain
                const uint32_t rcn = rl[(q + 1) * 32];
                const uint32_t x = rcw ^ c4;
                int grun = kJumpNever;
                int4 o;
                { const int keep = run; run = grun; o.x = cell(x, 0, g.x, 3); o.y = cell(x, 1, g.y, 2); o.z = cell(x, 2, g.z, 1); o.w = cell(x, 3, g.w, 0);
                  grun = run; run = keep; }
                if (!kJump) run = __viaddmax_s32(grun, kq, run);
                else lm = __vimax3_s32(__vimax3_s32(lm, o.x, o.y), o.z, o.w);      // matrix maximum, as G (two instructions per group)
                gl[q * 32] = o;
                kq -= 4;
                g = gn; rcw = rcn;
            }
            // last group: cells c >= dcap are phantoms; the state that leaves the strip is the one before cell dcap
            int left_cap = left, run_cap = run, lm_cap = lm;
            {
                const uint32_t x = rcw ^ c4;
                int4 o;
                o.x = cell(x, 0, g.x, kq + 3);
                if (dcap >= 1) { left_cap = left; run_cap = run; lm_cap = max(lm_cap, left); }
                o.y = cell(x, 1, g.y, kq + 2);
                if (dcap >= 2) { left_cap = left; run_cap = run; lm_cap = max(lm_cap, left); }
                o.z = cell(x, 2, g.z, kq + 1);
                if (dcap >= 3) { left_cap = left; run_cap = run; lm_cap = max(lm_cap, left); }
                o.w = cell(x, 3, g.w, kq);
                if (dcap >= 4) { left_cap = left; run_cap = run; lm_cap = max(lm_cap, left); }
                gl[(ng - 1) * 32] = o;
            }
            last_out = left_cap;
            if (!kJump) {
                const int v32 = run_cap & ~31;
                if (v32 > rbv) { rbv = v32; rbj = first + 31 - (run_cap & 31); }      // strict: the first column keeps a tie
                rbv_out = rbv; rbj_out = rbj;
                if (lane == 31) rowbest[i] = jpack(rbv >> kJumpScaleShift, (uint32_t)rbj);
                words[(size_t)t * 32 + lane] = make_uint2(wdel & ~wins, wins);
            } else {
                lmax = max(lmax, lm_cap + g32);                      // lm is kept as G = 32 H - 32 g
                const int hl = last_out + g32;                       // lane 31: H(i, m) x 32
                if (hl > lc_v) { lc_v = hl; lc_i = i; }              // strict: the first row keeps a tie
                words[(size_t)t * 32 + lane] = make_uint2(wj | (wdel & ~wins), wins | wj);
            }
            diag_keep = left_in;
        }
    }
    JumpMatrixOut out;
    out.h2max = 0; out.lastrow = jpack(0, 0); out.lastcol = jpack(0, 0);
    if (kJump) {
        // every lane now holds row n of its strip: first maximum of the last row
        unsigned long long lrow = jpack(0, 0);
        if (n >= 1)
            for (int k = 0; k < nreal; ++k) {
                const int gk = reinterpret_cast<const int*>(gl + (k >> 2) * 32)[k & 3];
                const unsigned long long w = jpack((gk + g32) >> kJumpScaleShift, (uint32_t)(first + k));
                lrow = w > lrow ? w : lrow;
            }
        out.h2max = junpack_value(warp_max_u64(jpack(lmax, 0))) >> kJumpScaleShift;
        out.lastrow = warp_max_u64(lrow);
        out.lastcol = jpack(__shfl_sync(0xffffffffu, lc_v, 31) >> kJumpScaleShift, (uint32_t)__shfl_sync(0xffffffffu, lc_i, 31));
    }
    __syncwarp();
    return out;
}

// (lane, cell) of region column j (1-based) under the strip layout of jump_matrix
__device__ __forceinline__ void jump_cell_of(int j, int S, int a, int& l, int& k)
{
    const int full = a * S;
    if (j <= full) { l = (j - 1) / S; k = (j - 1) - l * S; }
    else { const int jj = j - full - 1, q = jj / (S - 1); l = a + q; k = jj - q * (S - 1); }      // S >= 2 here
}

__device__ __forceinline__ void jump_warp_problem(const JumpWarpParams& p, const JumpWarpProblem& pr, int4* gs, uint32_t* rcs)
{
    const int lane = threadIdx.x & 31;
    const int n = (int)pr.n, m1 = (int)pr.m1, m2 = (int)pr.m2;
    const uint8_t* contig = p.bytes + pr.c_off;
    uint2* W1 = reinterpret_cast<uint2*>(p.work + pr.w1_off);
    uint2* W2 = reinterpret_cast<uint2*>(p.work + pr.w2_off);
    unsigned long long* rowbest = p.work + pr.row_off;
    unsigned long long* prefix = rowbest + (n + 1);
    int4* rowinfo = reinterpret_cast<int4*>(prefix + (n + 1));      // 2 (n + 1) words after an even offset: 16-byte aligned

    if (lane == 0) rowbest[0] = jpack(0, 0);
    if (m1 == 0) for (int i = 1 + lane; i <= n; i += 32) rowbest[i] = jpack(0, 0);
    if (m1 > 0) jump_matrix<false>(p, contig, p.bytes + pr.r1_off, n, m1, W1, rowbest, nullptr, gs, rcs);
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
        const double lim = (double)(1 << 24);             // the host admits |H| < 2^23: a jump beyond +-2^24 never matters
        for (int i = 1 + lane; i <= n; i += 32) {
            const unsigned long long w = __ldcg(prefix + i - 1);
            const int hm = junpack_value(w);
            int4 info = make_int4(kJumpNever, kJumpNever, kJumpNever, kJumpNever);      // no jump from this row
            if (hm >= p.minimum_score) {
                const uint32_t dist = (uint32_t)i - junpack_index(w);
                const double lg = dist < p.logtab_len ? p.logtab[dist] : log((double)dist);
                const double prod = __dmul_rn((double)p.jump_cost, lg);
                double jm = __dsub_rn((double)(hm + p.match), prod), jx = __dsub_rn((double)(hm - p.mismatch_penalty), prod);
                jm = fmin(fmax(jm, -lim), lim); jx = fmin(fmax(jx, -lim), lim);
                constexpr int kScale = 1 << kJumpScaleShift;
                info = make_int4((int)ceil(jm) * kScale, (int)jm * kScale, (int)ceil(jx) * kScale, (int)jx * kScale);
            }
            rowinfo[i] = info;
        }
        __threadfence();
        __syncwarp();
    }
    JumpMatrixOut o2;
    o2.h2max = 0; o2.lastrow = jpack(0, 0); o2.lastcol = jpack(0, 0);        // m2 == 0: no cells, all boundary
    if (m2 > 0) o2 = jump_matrix<true>(p, contig, p.bytes + pr.r2_off, n, m2, W2, rowbest, rowinfo, gs, rcs);
    __threadfence();
    __syncwarp();

    // ---- traceback, the whole warp in lock step (every lane holds the same walk state)
    JumpResult res;
    res.found = 0; res.score = o2.h2max;
    res.i1_start = res.i1_end = res.i2_start = res.i2_end = 0;
    res.j1_start = res.j1_end = res.j2_start = res.j2_end = 0;
    res.n_ops2 = res.n_ops1 = 0;
    int win_l = -1, win_i = 0;
    uint2 wv = make_uint2(0u, 0u);
    auto path_at = [&](const uint2* W, const int S, const int a, const int i, const int j) -> int {
        int l, k;
        jump_cell_of(j, S, a, l, k);
        if (l != win_l || i > win_i || win_i - i > 31) {       // lane q fetches the word of row i - q of strip l
            win_l = l; win_i = i;
            const int r = i - lane;
            wv = r >= 1 ? __ldcg(W + (size_t)(r + l - 1) * 32 + l) : make_uint2(0u, 0u);
        }
        const int src = win_i - i;
        const uint32_t lo = __shfl_sync(0xffffffffu, wv.x, src), hi = __shfl_sync(0xffffffffu, wv.y, src);
        const int pos = 4 * ((S + 3) / 4) - 1 - k;       // every step pushes the bits of whole groups of four cells
        return 1 + (int)((lo >> pos) & 1u) + 2 * (int)((hi >> pos) & 1u);
    };
    const int S1 = (m1 + 31) / 32, a1 = m1 - 32 * (S1 - 1), S2 = (m2 + 31) / 32, a2 = m2 - 32 * (S2 - 1);
    int ic, jc;
    if (junpack_value(o2.lastrow) >= junpack_value(o2.lastcol)) { ic = n; jc = (int)junpack_index(o2.lastrow); }
    else { ic = (int)junpack_index(o2.lastcol); jc = m2; }
    const int i2_end = ic, j2_end = jc;
    uint8_t* ops = p.ops + pr.ops_off;
    uint32_t n2 = 0, n1 = 0;
    int jumped = 0, h1_score = 0, i1_end = 0, j1_end = 0, i2_start = 0, j2_start = 0;
    while (ic > 0 && jc > 0) {
        const int path = path_at(W2, S2, a2, ic, jc);
        if (lane == 0) ops[n2] = (uint8_t)path;
        ++n2;
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
        win_l = -1;
        while (i > 0 && j > 0) {
            const int path = path_at(W1, S1, a1, i, j);
            if (lane == 0) ops[n2 + n1] = (uint8_t)path;
            ++n1;
            if (path == 1) { --i; --j; }
            else if (path == 2) { --i; }
            else { --j; }
        }
        res.found = 1; res.n_ops1 = n1;
        res.i1_start = i; res.i1_end = i1_end; res.i2_start = i2_start; res.i2_end = i2_end;
        res.j1_start = j + 1; res.j1_end = j1_end; res.j2_start = j2_start; res.j2_end = j2_end;
    }
    if (lane == 0) p.results[pr.res_idx] = res;
    __syncwarp();
}

__global__ void __launch_bounds__(kJumpWarpsPerBlock * 32, 4) sw_jump_warp_kernel(const JumpWarpParams p)
{
    extern __shared__ __align__(16) uint8_t jwsm[];
    uint8_t* const wsm = jwsm + (size_t)(threadIdx.x >> 5) * kJumpWarpSmem;
    int4* const gs = reinterpret_cast<int4*>(wsm);
    uint32_t* const rcs = reinterpret_cast<uint32_t*>(wsm + (size_t)kJumpGroups * 32 * 16);
    for (;;) {
        uint32_t pi = 0;
        if ((threadIdx.x & 31) == 0) pi = atomicAdd(p.next, 1u);
        pi = __shfl_sync(0xffffffffu, pi, 0);
        if (pi >= p.n_problems) break;
        const JumpWarpProblem pr = p.problems[pi];
        jump_warp_problem(p, pr, gs, rcs);
    }
}

}  // namespace nmsv
