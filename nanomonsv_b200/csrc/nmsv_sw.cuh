// nmsv_sw.cuh -- the Smith-Waterman wavefront kernel (sm_100a, DPX s16x2).
//
// Replaces the arithmetic of parasail.ssw as called at reference count_sread_by_alignment.py:148-149
// (Gotoh affine-gap local alignment; end cell = first column reaching the maximum, then smallest row).
//
// Mapping (DESIGN.md "Kernel 1"):
//   * one warp = one task = (row sequence A, row sequence B) x one column sequence.  A and B live in the low /
//     high 16-bit half of every register, so one DPX instruction advances two independent alignments that share
//     the read column: the template and its reverse complement (hot path), or A alone (begin pass).
//   * rows are cut into tiles of 32*R; inside a tile lane l owns the R consecutive rows l*R .. l*R+R-1 and keeps
//     their H (previous column) and E in registers.  Lane l works on column t-l at step t (anti-diagonal
//     wavefront); H and F of a strip's last row go to lane l+1 with two __shfl_up_sync per step.
//   * the substitution scores come from a per-warp shared-memory profile prof[code][row] (LDS.128, conflict
//     free) so the inner loop has no compares; the read is staged through a shared-memory ring.
//   * templates longer than one tile: the tiles of a task are JOBS that the warps of one block claim in order, so the
//     tiles of a task run side by side on different warps, each lagging its producer by a few 32-column groups; the H/F
//     row between two tiles is handed over through a small ring (1024 columns, 8 B each, L2 resident) with two
//     progress counters in shared memory: no scratch that grows with the read, no DRAM traffic, and the serial chain
//     of a task is one tile, not all of them. Only tasks with more tiles than a block has warps keep every fourth
//     boundary row in a buffer as long as the read.
//   * every stored value of (step t, strip row k) is  true value + a + (t - tbase + k) * ext  -- an anti-diagonal
//     bias: a value that moves one column to the right (E) or one row down (F) meets a bias that is ext larger,
//     so BOTH gap extensions cost no instruction; the zero floor of local alignment becomes a sliding window of
//     registers FL[u + k] computed on the FMA pipe. Per packed cell pair 4 DPX instructions on the ALU pipe:
//         x  = VIADDMNMX.S16x2(Hdiag, S + 2*ext, E)        max(Hdiag + S, E)
//         H  = VIMNMX3.S16x2  (x, FL, F)                   max(x, F, 0)
//         E' = VIADDMNMX.S16x2(H, ext - open, E)           max(H - open, E - ext)
//         F' = VIADDMNMX.S16x2(H, ext - open, F)           max(H - open, F - ext)
//     plus 1 DPX per 16 cells for the sampled running maximum. a (sw_bias) keeps the low half of every
//     value that is touched by a plain 32-bit add non-negative, so no borrow crosses the halves; the bias is
//     rebased at chunk boundaries before 16 bits could overflow.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nmsv {

constexpr int kPadCode = 4;        // columns outside the read use the wildcard code: it scores 0, so it can neither
                                   // start an alignment nor raise a maximum (updates need a strict increase)
constexpr int kRowPad = 6;         // row beyond the end of a template half: scores kPadScore against everything
constexpr int kNumCodes = 5;       // A C G T wildcard
constexpr int kPadScore = -128;
constexpr int kChunk = 128;        // columns staged per refill
constexpr int kRing = 4 * kChunk;  // read ring: the chunk in use, the previous one (lanes lag up to 31 columns) and
                                   // the next one, which TMA fills while the current chunk is computed
#ifndef NMSV_WPB
#define NMSV_WPB 4
#endif
constexpr int kWarpsPerBlock = NMSV_WPB;   // independent warps per block (each owns its slice of shared memory)
#ifndef NMSV_WARPS_PER_SM
#define NMSV_WARPS_PER_SM 16
#endif
constexpr unsigned kFull = 0xffffffffu;
#ifndef NMSV_SAMPLE_STRIDE
#define NMSV_SAMPLE_STRIDE 8       // the running maximum looks at every 8th row of a column (track()); measured on the
                                   // bench batch: stride 4 221.2 ms, 8 217.0 ms, 16 219.5 ms (fewer DPX vs more false alarms)
#endif
constexpr int kSnapWords = 2 * (32 / 4 + 1);   // per lane and half: up to 32 rows as uint4 + one word (maximum before, step)
#ifndef NMSV_UNROLL
#define NMSV_UNROLL 8
#endif
constexpr int kUnroll = NMSV_UNROLL;       // steps per iteration of the inner loop; measured 4 / 8 / 16: 217.1 / 210.1 ms /
                                           // register-bound (the floor window and the ring indices are shared by more steps)
static_assert(kUnroll == 4 || kUnroll == 8 || kUnroll == 16, "the 32-step groups are cut into kUnroll-step bodies");
constexpr int kMaxR = 32;          // largest strip height of any instantiation (head-room formula below)

// a: constant part of the bias
// (+1: the warp-wide lift subtracts 1 from a running maximum, which is never below a; the margin term keeps the
// carried threshold = maximum - margin non-negative from the first step on)
__host__ __device__ constexpr int sw_bias(int open, int ext) { return open + (NMSV_SAMPLE_STRIDE > 2 ? NMSV_SAMPLE_STRIDE - 1 : 1) * ext + 1; }
// 16-bit head-room a template needs on top of match * rows: the bias constant, one chunk of steps between two
// rebases plus the unrolled body, and the row part of the bias. The host rejects longer templates (NMSV_E_RANGE).
__host__ __device__ constexpr int sw_headroom(int open, int ext) { return sw_bias(open, ext) + ext + (kChunk + kUnroll + kMaxR) * ext; }
// Largest true H the 16-bit lanes are allowed to hold: parasail's own limit (NULL when a score exceeds
// INT16_MAX - match - 1) minus the head-room of the bias. A task whose scores cannot exceed match * min(rows, columns)
// <= this cap needs no run-time check; any other task is watched once per 32 steps (H grows by at most `match` per
// step) and reported as saturated (score -1, parasail: None) as soon as its maximum comes within 32 steps of the cap.
__host__ __device__ constexpr int sw_score_cap(int match, int open, int ext) { return 32767 - match - 1 - sw_headroom(open, ext); }

enum : uint32_t { kARev = 1u, kAComp = 2u, kBRev = 4u, kBComp = 8u, kCRev = 16u,
                  kTaskBegin = 32u };    // begin-pass task: result goes to rev_results and no continuation runs

struct SwTask {          // 64 bytes
    uint64_t a_base;     // index into codes of row 0 of half A (row i is a_base +/- i)
    uint64_t b_base;
    uint64_t c_base;     // index of column 0 (column j is c_base +/- j)
    uint32_t a_len;      // rows of half A (0 = unused half)
    uint32_t b_len;
    uint32_t ncols;      // 0 = empty task
    uint32_t flags;
    uint32_t out;        // index into results
    uint32_t rclass;     // strip height R of the kernel instantiation that owns this task
    uint32_t best0;      // half A: score the answer is known to have (begin pass: the forward score), 0 = unknown.
                         // The running maximum then starts at best0 - 1, so that only the cells that reach best0 --
                         // the candidates for the begin cell -- take the exact path instead of every cell of the
                         // climb along the alignment.
    uint32_t thr;        // forward task: only scores >= thr have to be exact (0 = all). The running maxima of both halves
                         // start at thr - 1: a half that never reaches thr reports thr - 1 with meaningless coordinates,
                         // and none of its cells takes the exact path of the position tracking.
    uint32_t hi;         // forward task: abandon the task as soon as a score reaches hi (0 = never): the caller only needs to
                         // know THAT it does (a reference allele that scores within the margin of the variant, :336-341);
                         // the result then reads score = hi for both halves, coordinates meaningless.
    uint32_t pad_[1];
};

static_assert(sizeof(SwTask) == 64, "the dynamic queue reads a task as four 16-byte words");

struct SwResult {        // [0] = half A, [1] = half B ; row/col 0-based in the task's own orientation
    int32_t score[2];    // -1: the 16-bit lanes could have saturated (parasail.ssw returns None)
    int32_t row[2];
    int32_t col[2];
};

// Continuation of a task group (validate path): when the last forward task of a read has finished, lane 0 of that warp
// calls sw_group_done(), which applies the decision rules to the read and appends its begin-pass tasks to a dynamic
// queue that the same persistent launch drains -- the begin pass fills the tail of the forward pass instead of running
// as an under-occupied launch of its own. Defined by the translation unit that includes this header.
struct GroupHook;
__device__ void sw_group_done(const GroupHook* hook, uint32_t out);

struct SwParams {
    const uint8_t* codes;
    const SwTask* tasks;
    const uint32_t* n_tasks_dev;   // when non-null the task count is read from device memory
    uint32_t n_tasks;
    uint32_t* queue_head;          // zeroed before the launch
    SwResult* fwd_results;         // results of forward tasks (static or queued by the continuation)
    SwResult* rev_results;         // results of begin-pass tasks (kTaskBegin)
    unsigned long long* cells;     // optional: [0] += forward cells executed, [1] += begin-pass cells executed
    uint2* bnd;                    // [grid blocks][kGBufs][bnd_stride] whole-read boundary buffers (tasks of > kW tiles only)
    uint32_t bnd_stride;
    uint2* rings;                  // [grid blocks][kRings][ring_cols] boundary rings between the tiles of a task
    uint32_t* ring_pos;            // [grid blocks][kRings] stream position of every ring (persists from launch to launch)
    uint32_t ring_cols;            // power of two
    uint32_t ring_shift;           // log2(ring_cols)
    uint4* snap;                   // [grid warps][kSnapWords * 32] column snapshots of the running-maximum tracker
    int match, mismatch, open, ext;
    uint32_t neg_open2;     // (-open, -open) packed: DPX add operand, read from the constant bank
    uint32_t ext_m_open2;   // (ext-open, ext-open) packed
    uint32_t ext2;          // ext * 65537
    // ---- optional continuation (all null/0 for independent tasks)
    const GroupHook* hook;          // device pointer
    uint32_t* group_remaining;      // [task.out >> 2] tasks of the group still to finish (set up before the launch)
    const SwTask* dyn_tasks;        // dynamic queue of this strip class: entries [0, *dyn_reserve), entry i valid once
    const uint32_t* dyn_ready;      //   dyn_ready[i] != 0; *dyn_head = next entry to hand out
    const uint32_t* dyn_reserve;
    uint32_t* dyn_head;
};

// ---- block-level job scheduler (shared memory) ------------------------------------------------------------------------
// A job = one tile (32*R rows) of one task over all columns. The warps of a block claim jobs in sequence (tile after tile
// of a task, task after task from the global lists); tile j+1 consumes the last row (H, F per column) of tile j through
// a link: a ring of ring_cols columns in global memory (it stays in L2: the rings of all blocks together are a few tens
// of MB that are rewritten every few hundred microseconds), or -- for the link into every kW-th job of a task with more
// than kW tiles, whose consumer can only be claimed a whole tile later -- one of two buffers as long as the read.
// A ring is one continuous stream of columns over the life of the context: link after link is appended without gaps,
// so every entry is written exactly once per lap, and each 8-byte entry carries the parity of its lap in a spare bit
// (H is never negative). The consumer therefore validates every entry it loads by itself -- no fence and no counter on
// the producer's side; the only counter is `consumed`, which keeps the producer from lapping the consumer.
// Whole-read buffers have no lap structure and use a `produced` counter behind a fence instead. (Rings in shared memory were measured first: 128
// columns is all that fits beside the profiles, and with so little slack the warps of a block -- which sit on different
// schedulers next to warps of other blocks and take the exact path of the running maximum at different times -- spent
// 9 % of their time waiting for each other.)
constexpr int kW = kWarpsPerBlock;
constexpr int kSlots = kW + 2;         // tasks in flight per block (at most kW running jobs + the task being dealt out)
constexpr int kRings = kW + 1;         // live ring links <= kW - 1 running consumers + the claimed job's in- and out-link
         // columns per ring; a consumer lags its producer by 2..5 groups of 32 columns
constexpr int kGBufs = 2;              // global boundary buffers per block
constexpr uint32_t kNoLink = 0xfu;
constexpr uint32_t kJobQuit = 0xffffffffu, kJobRetry = 0xfffffffeu;

struct alignas(16) Slot {              // one task in flight in this block (the task is read as 16-byte words)
    SwTask task;
    uint32_t ntiles, next_tile, tiles_done, sat;
    uint32_t lb[2];                    // per half: a score some cell of the task is known to reach (lower bound of the result)
    int32_t best_s[2], best_c[2], best_r[2];
    uint32_t last_out;                 // out-link of the tile claimed last = in-link of the next one
    uint32_t abort;                    // a tile saw a score >= task.hi: every tile of the task stops at its next chunk
    unsigned long long cells_done;     // cells the tiles of the task actually computed
};

// columns [0, produced) written (whole-read buffers only) / [0, consumed) no longer needed; base = position of the
// link's column 0 in the ring's stream
struct LinkCtl { uint32_t produced, consumed, base, users; };      // users: ends of the link (producer, consumer) not yet finished

struct BlockCtl {
    uint32_t lock, quit, n_waiting, cur;             // cur = slot whose tiles are being dealt out (kSlots = none)
    uint32_t slot_busy, ring_busy, gbuf_busy, seq;   // seq = sequence number of the next job
    uint32_t stream[kRings + kGBufs];                // per ring: stream position where its next link starts
    LinkCtl links[kRings + kGBufs];
    Slot slots[kSlots];
};
constexpr size_t kCtlBytes = (sizeof(BlockCtl) + 127) & ~(size_t)127;

template <int R>
__host__ __device__ constexpr int groups_of() { return (R + 3) / 4; }

template <int R>
__host__ __device__ constexpr size_t warp_smem_bytes()
{
    return (size_t)kNumCodes * groups_of<R>() * 32 * 16   // profile, uint4 [code][group][lane]
         + (size_t)kChunk * 8                               // boundary inputs of lane 0 for one chunk of columns, biased
         + (size_t)32 * 8                                   // boundary staging (outputs of lane 31), one group
         + (size_t)16                                       // constant slot read by lanes 1..31 instead of the ring
         + (size_t)16                                       // mbarrier of the TMA (cp.async.bulk) read staging
         + (size_t)2 * kRing;                               // read ring (4 chunk slots), stored twice (mirrored)
}

template <int R>
__host__ __device__ constexpr size_t block_smem_bytes()
{
    return kCtlBytes + (size_t)kWarpsPerBlock * warp_smem_bytes<R>();
}

__device__ __forceinline__ uint32_t pack16(int lo, int hi)
{
    return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu);
}

__device__ __forceinline__ int fetch_code(const uint8_t* __restrict__ codes, uint64_t base, int i, bool rev, bool comp)
{
    int c = codes[rev ? base - (uint64_t)i : base + (uint64_t)i];
    if (comp && c < 4) c = 3 - c;
    return c;
}

__device__ __forceinline__ int sub_score(int rc, int b, int match, int mismatch)
{
    if (rc == kRowPad) return kPadScore;
    if (rc > 3 || b > 3) return 0;
    return rc == b ? match : mismatch;
}

// x + k*65537 as one IMAD (FMA pipe): adds k to both 16-bit halves when no borrow/carry crosses the halves.
// Written as mad.lo so that ptxas keeps it off the ALU pipe, which the DPX instructions saturate.
__device__ __forceinline__ uint32_t add_halves(uint32_t x, int k)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 65537, %2;" : "=r"(r) : "r"((uint32_t)k), "r"(x));
    return r;
}

// ---- TMA (1-D bulk copy) staging of the read: global -> shared, completion on a per-warp mbarrier --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Potentially-blocking wait with a suspend-time hint. It is used WITHOUT a poll loop on purpose: a data-dependent
// loop (like a mutable per-thread "in flight" flag) makes the compiler drop its convergence / uniformity assumptions
// for the whole kernel (BRA.DIV before every shuffle, per-lane instead of uniform index arithmetic, +15 registers).
// The copy waited for was issued a whole chunk (~10^5 cycles) earlier.
__device__ __forceinline__ bool mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}

// bytes must be a multiple of 16, src and dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// lexicographic "better": higher score, then smaller column, then smaller row
__device__ __forceinline__ void warp_argbest(int& s, int& c, int& r)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        int os = __shfl_xor_sync(kFull, s, off);
        int oc = __shfl_xor_sync(kFull, c, off);
        int orw = __shfl_xor_sync(kFull, r, off);
        bool take = (os > s) || (os == s && (oc < c || (oc == c && orw < r)));
        if (take) { s = os; c = oc; r = orw; }
    }
}

// ---- links ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// A wait on another warp of the same block ends when that warp gets on with its tile; no legitimate wait comes near this
// bound (a whole 1 Mb read against a tile takes well under a second). Trapping turns a scheduling bug into an error
// return of the C ABI instead of a hung device.
constexpr unsigned long long kWatchdogNs = 20ull * 1000 * 1000 * 1000;

__device__ __noinline__ void link_wait_slow(const uint32_t* flag, uint32_t need, const uint32_t* abort)
{
    const unsigned long long t0 = global_ns();
    // warp-uniform exit (votes); an abandoned task ends the wait: the other side may be gone
    while (__any_sync(kFull, (int32_t)(ld_vol(flag) - need) < 0) && !__any_sync(kFull, ld_vol(abort) != 0u)) {
        __nanosleep(40);
        if (global_ns() - t0 > kWatchdogNs) __trap();
    }
}

// wait until *flag >= need (a column count written by another warp of this block). The common case -- the other side
// is far enough along -- is one shared-memory load and a vote; the poll loop is kept out of line (a data-dependent
// loop inlined into the kernel body costs the compiler its convergence analysis of the hot loop).
__device__ __forceinline__ void link_wait(const uint32_t* flag, uint32_t need, const uint32_t* abort)
{
    if (__any_sync(kFull, (int32_t)(ld_vol(flag) - need) < 0)) link_wait_slow(flag, need, abort);
    __threadfence_block();
}

// some ring entry of the warp is not yet written in this lap: poll (out of line, like every data-dependent loop, and
// called by the whole warp so that the call site stays convergent; lanes with stale = false return at once)
__device__ __noinline__ uint2 ring_poll(const uint2* ptr, uint32_t tag, uint2 v, bool stale, const uint32_t* abort)
{
    const unsigned long long t0 = global_ns();
    while (__any_sync(kFull, stale) && !__any_sync(kFull, ld_vol(abort) != 0u)) {        // warp-uniform exit
        __nanosleep(100);
        if (stale) {
            v = __ldcg(ptr);
            stale = ((v.x >> 15) & 1u) != tag;
        }
        if (global_ns() - t0 > kWatchdogNs) __trap();
    }
    return v;
}

__device__ __forceinline__ void ctl_lock(BlockCtl* c)
{
    const unsigned long long t0 = global_ns();
    while (atomicCAS(&c->lock, 0u, 1u) != 0u) {
        __nanosleep(20);
        if (global_ns() - t0 > kWatchdogNs) __trap();
    }
    __threadfence_block();
}

__device__ __forceinline__ void ctl_unlock(BlockCtl* c)
{
    __threadfence_block();
    atomicExch(&c->lock, 0u);
}

// Next task of this launch for the calling block (lane 0 of some warp, holding the block lock): the dynamic queue first
// (tasks queued by continuations: reference-allele alignments and begin passes of reads that are already decided), then
// the static list. Returns false when both are empty right now.
__device__ __forceinline__ bool pick_task(const SwParams& p, uint32_t n_tasks, SwTask& task)
{
    const SwTask* at = nullptr;
    if (p.dyn_reserve) {
        uint32_t h = ld_vol(p.dyn_head);
        while (h < ld_vol(p.dyn_reserve)) {
            const uint32_t o = atomicCAS(p.dyn_head, h, h + 1u);
            if (o == h) { at = p.dyn_tasks + h; break; }
            h = o;
        }
        // the producer publishes dyn_ready[h] right after the task body (fence in between); it is a running warp
        // between two adjacent statements, so this wait is a few hundred nanoseconds at most
        if (at) {
            while (ld_vol(p.dyn_ready + h) == 0u) __nanosleep(64);
            __threadfence();
        }
    }
    if (!at && ld_vol(p.queue_head) < n_tasks) {      // (checked first: idle warps poll, the counter must not wrap)
        const uint32_t ti = atomicAdd(p.queue_head, 1u);
        if (ti < n_tasks) at = p.tasks + ti;
    }
    if (!at) return false;
    // through L2 (a line of the dynamic queue may be stale in L1)
    const uint4* q = reinterpret_cast<const uint4*>(at);
    const uint4 w0 = __ldcg(q), w1 = __ldcg(q + 1), w2 = __ldcg(q + 2), w3 = __ldcg(q + 3);
    task.a_base = ((uint64_t)w0.y << 32) | w0.x; task.b_base = ((uint64_t)w0.w << 32) | w0.z;
    task.c_base = ((uint64_t)w1.y << 32) | w1.x; task.a_len = w1.z; task.b_len = w1.w;
    task.ncols = w2.x; task.flags = w2.y; task.out = w2.z; task.rclass = w2.w; task.best0 = w3.x; task.thr = w3.y; task.hi = w3.z;
    return true;
}

// Claim the next job of this block (lane 0). Returns slot | tile << 4 | in-link << 24 | out-link << 28, or kJobQuit.
// A warp that finds nothing to claim waits while another warp of its block is still running (the continuation of that
// warp's task may queue new tasks, and the tiles of a multi-tile task need all the warps of the block); the block
// leaves when all its warps wait at once and both lists are empty. Tasks queued later by other blocks are run by
// those blocks themselves, or by the follow-up launches over the same queue.
__device__ __noinline__ uint32_t claim_job(BlockCtl* ctl, const SwParams& p, uint32_t n_tasks, uint32_t R)
{
    bool counted = false;
    for (;;) {
        ctl_lock(ctl);
        uint32_t ret = kJobRetry;
        if (ctl->quit) ret = kJobQuit;
        else {
            if (ctl->cur == (uint32_t)kSlots) {
                SwTask task;
                task.pad_[0] = 0;
                if (pick_task(p, n_tasks, task)) {
                    if (task.rclass == R && task.ncols != 0) {
                        const uint32_t si = (uint32_t)__ffs((int)~ctl->slot_busy) - 1u;      // a free slot exists (kSlots > kW)
                        ctl->slot_busy |= 1u << si;
                        Slot& s = ctl->slots[si];
                        s.task = task;
                        const uint32_t rows = task.a_len > task.b_len ? task.a_len : task.b_len;
                        s.ntiles = (rows + 32u * R - 1u) / (32u * R);
                        s.next_tile = 0; s.tiles_done = 0; s.sat = 0; s.abort = 0; s.cells_done = 0;
                        // begin pass: the score is known (only cells that reach it matter); its unused half B holds the
                        // floor everywhere, park its bound out of reach so that it never takes the exact path
                        s.lb[0] = task.best0 > task.thr ? task.best0 : task.thr; s.lb[1] = task.b_len ? task.thr : 256u;
                        s.best_s[0] = (int32_t)task.best0; s.best_c[0] = task.best0 ? 0x7fffffff : 0;
                        s.best_r[0] = task.best0 ? 0x7fffffff : (int32_t)task.a_len - 1;
                        s.best_s[1] = 0; s.best_c[1] = 0; s.best_r[1] = (int32_t)task.b_len - 1;
                        s.last_out = kNoLink;
                        ctl->cur = si;
                        // ring links never cross a multiple of kW in the job sequence: a task of up to kW tiles is not
                        // split over such a boundary, longer ones start on one
                        const uint32_t ph = ctl->seq % (uint32_t)kW;
                        if (ph && (s.ntiles > (uint32_t)kW || ph + s.ntiles > (uint32_t)kW)) ctl->seq += (uint32_t)kW - ph;
                    }
                    // (a task of another strip class, or an empty one, is skipped; look again without sleeping)
                    else { ctl_unlock(ctl); continue; }
                } else {
                    if (!counted) { counted = true; ctl->n_waiting += 1u; }
                    if (ctl->n_waiting == (uint32_t)kW) { ctl->quit = 1u; ret = kJobQuit; }
                }
            }
            if (ctl->cur != (uint32_t)kSlots) {
                Slot& s = ctl->slots[ctl->cur];
                const uint32_t j = s.next_tile;
                uint32_t out = kNoLink;
                bool ok = true;
                if (j + 1u < s.ntiles) {
                    if ((ctl->seq + 1u) % (uint32_t)kW == 0u) {          // the consumer is claimed a tile later: global link
                        const uint32_t freeg = ~ctl->gbuf_busy & ((1u << kGBufs) - 1u);
                        if (freeg) { out = (uint32_t)kRings + (uint32_t)__ffs((int)freeg) - 1u; ctl->gbuf_busy |= 1u << (out - kRings); }
                        else ok = false;                                 // both in use: their consumers are running, retry
                    } else {
                        // kRings = kW + 1 rings cover every link with a running end in normal operation; while abandoned
                        // tasks wind down (a consumer gone, its producer still on its last chunk) there can be more: retry
                        const uint32_t freer = ~ctl->ring_busy;
                        if (freer) { out = (uint32_t)__ffs((int)freer) - 1u; ctl->ring_busy |= 1u << out; }
                        else ok = false;
                    }
                }
                if (ok) {
                    ret = ctl->cur | (j << 4) | (s.last_out << 24) | (out << 28);
                    if (out != kNoLink) {
                        ctl->links[out].produced = 0u; ctl->links[out].consumed = 0u; ctl->links[out].users = 2u;
                        ctl->links[out].base = ctl->stream[out];
                        ctl->stream[out] += s.task.ncols;
                    }
                    s.last_out = out;
                    s.next_tile = j + 1u;
                    ctl->seq += 1u;
                    if (s.next_tile == s.ntiles) ctl->cur = (uint32_t)kSlots;
                    if (counted) { counted = false; ctl->n_waiting -= 1u; }
                }
            }
        }
        ctl_unlock(ctl);
        if (ret != kJobRetry) return ret;
        __nanosleep(500);
    }
}

// lane 0, after the result of a static task is stored: count the task's group down and run the continuation when
// this was its last task
__device__ __noinline__ void sw_task_done(const GroupHook* hook, uint32_t* group_remaining, uint32_t out)
{
    __threadfence();        // the result before the count
    if (atomicSub(&group_remaining[out >> 2], 1u) == 1u) {
        __threadfence();    // the other tasks' results after their counts
        sw_group_done(hook, out);
    }
}

// An abandoned task (Slot::abort): the producer of a link still owes the ring the rest of the link's stream positions --
// every ring entry has to be written once per lap, with the parity of its lap, or a later link would take a stale entry
// for a fresh one. Only the last lap of the range decides what the entries hold in the end. Whole-read buffers just
// publish the full count. The consumer is abandoning the task too and ignores what it reads.
__device__ __noinline__ void link_abort_fill(const SwParams& p, uint2* base, uint32_t mask, bool ring, uint32_t q0, int from_col,
                                             int ncols, LinkCtl* link)
{
    const int lane = threadIdx.x & 31;
    if (ring) {
        int first = ncols - (int)p.ring_cols;
        if (first < from_col) first = from_col;
        for (int col = first + lane; col < ncols; col += 32) {
            const uint32_t q = q0 + (uint32_t)col;
            __stcg(base + (q & mask), make_uint2((((q >> p.ring_shift) + 1u) & 1u) << 15, 0u));
        }
    } else {
        __threadfence();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&link->produced) = (uint32_t)ncols;
    }
    __syncwarp();
}

// lane 0, at the end of a tile: merge the tile's best cells into the task's result (higher score, then smaller column,
// then smaller row -- parasail's end cell), release the tile's in-link, and when this was the task's last tile to finish
// store the result and run the continuation.
__device__ __noinline__ void finish_job(BlockCtl* ctl, const SwParams& p, uint32_t code, int s0, int c0, int r0,
                                        int s1, int c1, int r1, uint32_t sat, unsigned long long cells_done)
{
    const uint32_t si = code & 15u, in_link = (code >> 24) & 15u, out_link = code >> 28;
    Slot& s = ctl->slots[si];
    ctl_lock(ctl);
    if (s0 > 0 && (s0 > s.best_s[0] || (s0 == s.best_s[0] && (c0 < s.best_c[0] || (c0 == s.best_c[0] && r0 < s.best_r[0]))))) {
        s.best_s[0] = s0; s.best_c[0] = c0; s.best_r[0] = r0;
    }
    if (s1 > 0 && (s1 > s.best_s[1] || (s1 == s.best_s[1] && (c1 < s.best_c[1] || (c1 == s.best_c[1] && r1 < s.best_r[1]))))) {
        s.best_s[1] = s1; s.best_c[1] = c1; s.best_r[1] = r1;
    }
    s.sat |= sat;
    s.cells_done += cells_done;
    // a link is handed out again when both of its ends are done with it (the producer of an abandoned task may still be
    // filling the ring when its consumer has already left)
    for (int e = 0; e < 2; ++e) {
        const uint32_t l = e ? out_link : in_link;
        if (l == kNoLink) continue;
        if (--ctl->links[l].users == 0u) {
            if (l < (uint32_t)kRings) ctl->ring_busy &= ~(1u << l);
            else ctl->gbuf_busy &= ~(1u << (l - (uint32_t)kRings));
        }
    }
    s.tiles_done += 1u;
    const bool done = s.tiles_done == s.ntiles;
    SwResult res;
    uint32_t out = 0, flags = 0;
    unsigned long long cells = 0;
    if (done) {
        res.score[0] = (s.sat & 1u) ? -1 : s.best_s[0]; res.row[0] = s.best_r[0]; res.col[0] = s.best_c[0];
        res.score[1] = s.task.b_len ? ((s.sat & 2u) ? -1 : s.best_s[1]) : 0; res.row[1] = s.best_r[1]; res.col[1] = s.best_c[1];
        if (s.abort) {       // abandoned: all the caller learns is that a score reaches hi
            res.score[0] = (int32_t)s.task.hi; res.score[1] = s.task.b_len ? (int32_t)s.task.hi : 0;
            res.row[0] = res.row[1] = res.col[0] = res.col[1] = 0;
        }
        out = s.task.out; flags = s.task.flags;
        cells = s.cells_done;
        ctl->slot_busy &= ~(1u << si);
    }
    ctl_unlock(ctl);
    if (!done) return;
    const bool is_begin = flags & kTaskBegin;
    (is_begin ? p.rev_results : p.fwd_results)[out] = res;
    if (p.cells) atomicAdd(&p.cells[is_begin ? 1 : 0], cells);
    if (!is_begin && p.hook) sw_task_done(p.hook, p.group_remaining, out);
}


// OPEN/EXT >= 0: gap penalties known at compile time (the reference always calls parasail.ssw(.., 3, 1, ..),
// count_sread_by_alignment.py:148): the DPX add operands become immediates, which saves one vector-register read on
// two of the four DPX instructions per cell pair. OPEN = EXT = -1: generic instantiation, operands from the parameters.
template <int R, int OPEN, int EXT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, ((R <= 16 ? NMSV_WARPS_PER_SM : (R <= 19 ? 12 : 8)) / kWarpsPerBlock > 0 ? (R <= 16 ? NMSV_WARPS_PER_SM : (R <= 19 ? 12 : 8)) / kWarpsPerBlock : 1))
sw_wavefront_kernel(const SwParams p)
{
    constexpr int G = groups_of<R>();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    BlockCtl* const ctl = reinterpret_cast<BlockCtl*>(smem_raw);
    uint2* const rings = p.rings + (size_t)blockIdx.x * kRings * p.ring_cols;          // [kRings][ring_cols]
    uint8_t* wbase = smem_raw + kCtlBytes + (size_t)warp * warp_smem_bytes<R>();
    uint4* prof = reinterpret_cast<uint4*>(wbase);
    uint2* bring = reinterpret_cast<uint2*>(wbase + (size_t)kNumCodes * G * 32 * 16);
    uint2* bstage = bring + kChunk;
    uint2* bconst = bstage + 32;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(bconst + 2);
    uint8_t* rring = reinterpret_cast<uint8_t*>(bconst + 4);
    const uint32_t gwarp = blockIdx.x * kWarpsPerBlock + warp;
    uint2* const gbuf0 = p.bnd + (size_t)blockIdx.x * kGBufs * p.bnd_stride;          // [kGBufs][bnd_stride]
    uint4* const snap_lane = p.snap + (size_t)gwarp * (kSnapWords * 32) + lane;    // word w of half h: [(h * (G+1) + w) * 32]

    const uint32_t n_tasks = p.n_tasks_dev ? *p.n_tasks_dev : p.n_tasks;
    // Every stored H/E/F value of (step t, strip row k) is  true value + a + (t - tbase + k) * ext  (header comment):
    //   * the bias grows by ext per step AND per row, which turns  E' = max(H - open, E - ext)  and
    //     F' = max(H - open, F - ext)  into  max(H + (ext - open), E / F): neither gap extension costs an instruction;
    //   * a = sw_bias keeps the low half of everything that a plain 32-bit add touches non-negative (no borrow between
    //     the halves); the step part is rebased at chunk boundaries.
    const int bias = sw_bias(p.open, p.ext);
    const uint32_t FLOOR0 = pack16(bias, bias);
    const uint32_t EF0 = pack16(bias - p.open, bias - p.open);          // E = F = -open
    constexpr bool kFixedGaps = (OPEN >= 0 && EXT >= 0);
    const uint32_t EXT_M_OPEN = kFixedGaps ? (((uint32_t)(EXT - OPEN) & 0xffffu) * 65537u) : p.ext_m_open2;   // H + ext - open
    const int ext = p.ext;
    const bool is_lane0 = (lane == 0);
    // lane 0 takes H/F of the row above its strip from the boundary ring, lanes 1..31 from lane-1 by shuffle. Instead
    // of two SELs (ALU pipe) per step every lane does  in = shfl * m1 + ld  (IMAD): lane 0 has m1 = 0 and loads the
    // ring entry of this step, the others have m1 = 1 and load a constant slot holding +ext (the bias lift).
    uint32_t m1 = is_lane0 ? 0u : 1u;
    asm volatile("" : "+r"(m1));           // opaque: keeps ptxas from re-deriving it from the lane id inside the loop
    const char* const bsrc = is_lane0 ? reinterpret_cast<const char*>(bring) : reinterpret_cast<const char*>(bconst);
    uint32_t bstride = is_lane0 ? 8u : 0u;
    asm volatile("" : "+r"(bstride));      // opaque: otherwise ptxas specialises the address arithmetic with predicated moves
    // read ring: column j lives at rring[j % kRing] and again kRing bytes later; lane l reads (t % kRing) + kRing - l
    const uint8_t* const rlane = rring + kRing - lane;
    // ext once more, in a vector register: as a uniform value ptxas re-materialises it with a move per use inside the
    // hot loop. A value that went through shared memory is not uniform in its eyes.
    if (lane == 0) bconst[1] = make_uint2((uint32_t)p.ext, 0u);
    __syncwarp();
    const uint32_t ext_v = reinterpret_cast<const volatile uint2*>(bconst)[1].x;
    if (lane == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        ctl->lock = 0; ctl->quit = 0; ctl->n_waiting = 0; ctl->cur = (uint32_t)kSlots;
        ctl->slot_busy = ~((1u << kSlots) - 1u); ctl->ring_busy = ~((1u << kRings) - 1u); ctl->gbuf_busy = 0; ctl->seq = 0;
        for (int r = 0; r < kRings + kGBufs; ++r) ctl->stream[r] = r < kRings ? p.ring_pos[blockIdx.x * kRings + r] : 0u;
    }
    __syncthreads();               // the only block-wide barrier: from here on the warps meet through the links only
    uint32_t tma_parity = 0;       // phase of the next mbarrier completion to wait for

    for (;;) {
        // ---- next job: a tile of the task this block is dealing out, or the first tile of a new task
        uint32_t code = 0;
        if (lane == 0) code = claim_job(ctl, p, n_tasks, (uint32_t)R);
        code = __shfl_sync(kFull, code, 0);
        if (code == kJobQuit) break;
        Slot* const slot = ctl->slots + (code & 15u);
        const int tile = (int)((code >> 4) & 0xfffffu);
        const uint32_t in_link = (code >> 24) & 15u, out_link = code >> 28;
        SwTask task;
        {
            const uint4* q = reinterpret_cast<const uint4*>(&slot->task);
            const uint4 w0 = q[0], w1 = q[1], w2 = q[2];
            task.a_base = ((uint64_t)w0.y << 32) | w0.x; task.b_base = ((uint64_t)w0.w << 32) | w0.z;
            task.c_base = ((uint64_t)w1.y << 32) | w1.x; task.a_len = w1.z; task.b_len = w1.w;
            task.ncols = w2.x; task.flags = w2.y; task.out = w2.z; task.rclass = w2.w;
            task.hi = reinterpret_cast<const uint32_t*>(&slot->task.hi)[0];
        }

        const int a_len = (int)task.a_len, b_len = (int)task.b_len, ncols = (int)task.ncols;
        const bool a_rev = task.flags & kARev, a_comp = task.flags & kAComp;
        const bool b_rev = task.flags & kBRev, b_comp = task.flags & kBComp;
        const bool c_rev = task.flags & kCRev;
        const int rows = a_len > b_len ? a_len : b_len;
        const int total_steps = (ncols + 31 + kUnroll - 1) & ~(kUnroll - 1);   // the step body is unrolled kUnroll x (pad steps are inert)
        // steps the bias may run before the 16-bit lanes could overflow (>= kChunk by construction of the cap)
        const int hbound = p.match * (rows < ncols ? rows : ncols);       // no alignment of this task can score more
        const int hcap = sw_score_cap(p.match, p.open, ext);
        const int hmax = hbound < hcap ? hbound : hcap;
        const int budget_steps = ext > 0 ? (32767 - bias - ext - hmax) / ext - kMaxR - kUnroll : 0x7fffffff;
        // run-time saturation watch (only tasks whose bound exceeds the cap can trip it)
        const uint32_t sat_thr = hbound <= hcap ? 0xffffu : (uint32_t)(hcap - 32 * p.match);
        uint32_t sat = 0;          // bit h: half h came within 32 steps of the cap
        const uint32_t hi = task.hi;          // abandon the task when a score reaches hi (0: never)
        int cols_done = ncols;                // columns this tile went through (fewer when the task was abandoned)
        int left_at = -1;                     // step at which an abandoned task was left (-1: ran to the end)

        {
            const int row0 = tile * 32 * R + lane * R;
            // Links: the row above this tile comes in column by column from the warp that runs tile-1 (in_link), this
            // tile's last row goes out to the warp that runs tile+1 (out_link). Ids below kRings are shared-memory
            // rings, the others global buffers (updated in place: a later producer of the same buffer is a dependent of
            // its earlier consumer, and the scheduler hands a buffer out again only when its consumer has finished).
            const bool has_in = in_link != kNoLink, has_out = out_link != kNoLink;
            const bool in_ring = in_link < (uint32_t)kRings, out_ring = out_link < (uint32_t)kRings;
            // column c of a link lives at base[c & mask]: a ring wraps, a whole-read buffer does not
            const uint2* const b_in = in_ring ? rings + (size_t)in_link * p.ring_cols
                                              : gbuf0 + (size_t)(has_in ? in_link - kRings : 0u) * p.bnd_stride;
            uint2* const b_out = out_ring ? rings + (size_t)out_link * p.ring_cols
                                          : gbuf0 + (size_t)(has_out ? out_link - kRings : 0u) * p.bnd_stride;
            const uint32_t m_in = in_ring ? p.ring_cols - 1u : 0xffffffffu, m_out = out_ring ? p.ring_cols - 1u : 0xffffffffu;
            LinkCtl* const lin = ctl->links + (has_in ? in_link : 0u);
            LinkCtl* const lout = ctl->links + (has_out ? out_link : 0u);
            const uint32_t q_in = in_ring ? lin->base : 0u, q_out = out_ring ? lout->base : 0u;     // stream positions of column 0
            const bool stage_lane = has_out && lane == 31;

            __syncwarp();
            // ---- profile of this tile: prof[code][group][lane].{x,y,z,w} = (S_A | S_B << 16) + 2*ext of row 4*group+w
            {
                uint32_t* pw = reinterpret_cast<uint32_t*>(prof);
                for (int k = 0; k < 4 * G; ++k) {
                    const int row = row0 + k;
                    const int ca = (k < R && row < a_len) ? fetch_code(p.codes, task.a_base, row, a_rev, a_comp) : kRowPad;
                    const int cb = (k < R && row < b_len) ? fetch_code(p.codes, task.b_base, row, b_rev, b_comp) : kRowPad;
#pragma unroll
                    for (int b = 0; b < kNumCodes; ++b) {
                        const uint32_t w = pack16(sub_score(ca, b, p.match, p.mismatch) + 2 * p.ext,
                                                  sub_score(cb, b, p.match, p.mismatch) + 2 * p.ext);
                        pw[(((b * G) + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = w;
                    }
                }
#pragma unroll
                for (int u = 0; u < (2 * kRing) / 32; ++u) rring[u * 32 + lane] = (uint8_t)kPadCode;
                if (lane == 0) {      // bias change from (step t-1, row R-1 / R) of the lane above to (step t, row -1 / 0)
                    const uint32_t lift = (uint32_t)(-((R - 1) * ext * 65537));
                    bconst[0] = make_uint2(lift, lift);
                }
            }

            // H of the previous column ping-pongs between HA and HB over the unrolled body, so that no register moves
            // are needed to keep Hdiag alive while the new column is written.
            uint32_t HA[R], HB[R], E[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {       // zero state with the bias of "after step -1": row k carries k*ext
                HA[k] = FLOOR0 + (uint32_t)k * p.ext2; HB[k] = HA[k]; E[k] = EF0 + (uint32_t)(k + 1) * p.ext2;
            }
            // Filter threshold for TWO consecutive columns at once: an unsampled row lies at most stride-1 rows above a
            // sampled one and exceeds its H by at most open + (stride-2)*ext; the older column is compared at the newer
            // column's bias, i.e. one ext too strictly, hence open + (stride-1)*ext.
            static_assert(NMSV_SAMPLE_STRIDE >= 2, "the threshold must stay non-negative: sw_bias + 2*ext >= margin");
            const uint32_t MARGIN2 = (uint32_t)((p.open + (NMSV_SAMPLE_STRIDE - 1) * ext) * 65537);
            // the running maximum starts at (a score the task is already known to reach) - 1, per half: only cells that
            // can still win or tie are examined. The bound is shared by the tiles of the task through the slot.
            uint32_t bestprev;
            {
                const int l0 = (int)ld_vol(&slot->lb[0]), l1 = (int)ld_vol(&slot->lb[1]);
                bestprev = FLOOR0 + pack16(l0 > 0 ? l0 - 1 : 0, l1 > 0 ? l1 - 1 : 0);
            }
            uint32_t flb = FLOOR0 + p.ext2;     // floor of row k at step t is flb + (t+k)*ext2 = FLOOR0 + (t - tbase + k)*ext2
            uint32_t h_last = FLOOR0 + (uint32_t)(R - 1) * p.ext2, f_out = EF0 + (uint32_t)R * p.ext2, diag = FLOOR0 - p.ext2;
            int tbase = -1;          // bias of row k at step t = (t - tbase + k) * ext ; "after step -1" row 0 carries 0
            // TMA staging applies to forward reads that start on a 16-byte boundary; whether a given chunk was
            // prefetched is a pure function of t0 (a mutable flag would cost the compiler its uniformity analysis)
            const bool tma_task = !c_rev && ((task.c_base & 15ull) == 0);

            // One column. tq = t - U is a multiple of 4, so every ring index of step t is a base computed once per four
            // steps plus the compile-time constant U (no per-step uniform-datapath arithmetic).
            auto column = [&](const uint32_t (&Hs)[R], uint32_t (&Hd)[R], const uint32_t (&FL)[R + kUnroll - 1], const char* bq,
                              const uint8_t* rq, uint2* sq, const int U) {
                const uint32_t sh = __shfl_up_sync(kFull, h_last, 1);
                const uint32_t sf = __shfl_up_sync(kFull, f_out, 1);
                const uint2 bv = *reinterpret_cast<const uint2*>(bq + (uint32_t)U * bstride);
                uint32_t hin, fin;     // H / F of the row above this strip, lifted to this step's bias
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hin) : "r"(sh), "r"(m1), "r"(bv.x));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(fin) : "r"(sf), "r"(m1), "r"(bv.y));
                const uint32_t b = rq[U];
                uint32_t S[4 * G];
                const uint4* pp = prof + b * (G * 32) + lane;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const uint4 v = pp[g * 32];
                    S[4 * g + 0] = v.x; S[4 * g + 1] = v.y; S[4 * g + 2] = v.z; S[4 * g + 3] = v.w;
                }
                uint32_t F = fin;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const uint32_t x = __viaddmax_s16x2(k == 0 ? diag : Hs[k - 1], S[k], E[k]);  // max(Hdiag + S, E)
                    const uint32_t h = __vimax3_s16x2(x, FL[U + k], F);                          // H = max(x, F, 0)
                    E[k] = __viaddmax_s16x2(h, EXT_M_OPEN, E[k]);                                // E' = max(H-open, E-ext)
                    F = __viaddmax_s16x2(h, EXT_M_OPEN, F);                                      // F' = max(H-open, F-ext)
                    Hd[k] = h;
                }
                diag = hin;
                h_last = Hd[R - 1];
                f_out = F;
                if (stage_lane) sq[U] = make_uint2(h_last, f_out);
            };
            // ---- running maximum after a pair of columns (older column in Ha, newer in Hb; t = step of the older).
            // Rows k and k' <= k of one column satisfy H[k] >= H[k'] - open - (k-k'-1)*ext (through F), so if no row
            // of {R-1, R-1-stride, ...} comes within the margin of the running maximum no row of the strip exceeds it:
            // 1 DPX per `stride` rows and column in the common case.
            auto negk = [&](const int k) -> uint32_t {      // (-k*ext, -k*ext): removes the row part of the bias
                return kFixedGaps ? (((uint32_t)(-k * EXT)) & 0xffffu) * 65537u : (((uint32_t)(-k * ext)) & 0xffffu) * 65537u;
            };
            auto snap_store = [&](const int h, const uint32_t (&Hx)[R], const uint32_t before, const int step) {
                uint4* d = snap_lane + h * (G + 1) * 32;
#pragma unroll
                for (int g = 0; g < G; ++g)
                    __stcg(d + g * 32, make_uint4(Hx[4 * g], 4 * g + 1 < R ? Hx[(4 * g + 1) % R] : 0u,
                                                  4 * g + 2 < R ? Hx[(4 * g + 2) % R] : 0u, 4 * g + 3 < R ? Hx[(4 * g + 3) % R] : 0u));
                __stcg(d + G * 32, make_uint4(before, (uint32_t)step, 0u, 0u));
            };
            // (column, row) of the first cell of the saved column that holds its maximum, for half h
            auto resolve_snap = [&](const int h, int& col, int& row) {
                const uint4* d = snap_lane + h * (G + 1) * 32;
                const uint4 meta = __ldcg(d + G * 32);
                const int sh = 16 * h;
                int bestv = (int)(int16_t)(meta.x >> sh), r = 0;
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const uint4 v = __ldcg(d + g * 32);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = 4 * g + i;
                        const int hv = (int)(int16_t)(w[i] >> sh) - k * ext;      // row part of the bias removed
                        if (k < R && hv > bestv) { bestv = hv; r = k; }
                    }
                }
                col = (int)meta.y - lane;
                row = row0 + r;
            };
            auto track = [&](const uint32_t (&Ha)[R], const uint32_t (&Hb)[R], const int t) {
                const uint32_t bp2 = add_halves(bestprev, 2 * ext);      // running maximum at the newer step's row-0 bias
                const uint32_t thr = bp2 - MARGIN2;                       // bp2 >= a + 2*ext >= margin in both halves: no borrow
                uint32_t q = thr;
#pragma unroll
                for (int k = R - 1; k >= 0; k -= NMSV_SAMPLE_STRIDE) {
                    q = __viaddmax_s16x2(Ha[k], negk(k), q);
                    q = __viaddmax_s16x2(Hb[k], negk(k), q);
                }
                uint32_t best = bp2;
                if (q != thr) {
                    // rare: some row may exceed this lane's running maximum. Only the LAST rise of a lane can be the
                    // answer, and most rises are followed by another one a column later (the climb along the
                    // alignment), so the position is not worked out here: the exact new maximum is computed (one DPX per
                    // row), and for each half that rose the column it rose in is saved to a per-lane scratch slot together
                    // with the maximum before that column and the step; the slot is examined once, at the end of the
                    // tile (resolve_snap). A later tie does not rise and leaves the slot alone: first column wins.
                    const uint32_t b0 = add_halves(bestprev, ext);        // running maximum at (older step, row 0)
                    uint32_t ma = b0;
#pragma unroll
                    for (int k = 0; k < R; ++k) ma = __viaddmax_s16x2(Ha[k], negk(k), ma);
                    const uint32_t mb0 = add_halves(ma, ext);             // ... at (newer step, row 0)
                    uint32_t mb = mb0;
#pragma unroll
                    for (int k = 0; k < R; ++k) mb = __viaddmax_s16x2(Hb[k], negk(k), mb);
                    const uint32_t ra = ma ^ b0, rb = mb ^ mb0;           // non-zero half: that half rose in that column
                    if (rb & 0xffffu) snap_store(0, Hb, mb0, t + 1);
                    else if (ra & 0xffffu) snap_store(0, Ha, b0, t);
                    if (rb >> 16) snap_store(1, Hb, mb0, t + 1);
                    else if (ra >> 16) snap_store(1, Ha, b0, t);
                    best = mb;
                }
                bestprev = best;
            };

            for (int t0 = 0; t0 < total_steps; t0 += kChunk) {
                // ---- rebase the bias when the next chunk could leave the 16-bit budget (warp-uniform, rare)
                if (t0 + kChunk - tbase > budget_steps) {
                    const uint32_t d = (uint32_t)((t0 - 1 - tbase) * ext) * 65537u;
#pragma unroll
                    for (int k = 0; k < R; ++k) { HA[k] -= d; E[k] -= d; }
                    bestprev -= d; h_last -= d; f_out -= d; diag -= d;
                    tbase = t0 - 1;
                    flb = FLOOR0 - (uint32_t)tbase * p.ext2;
                }
                // ---- refill: columns [t0, t0+kChunk) of the read ring and of lane 0's boundary inputs (the last row of the
                //      tile above, from its link once the producer has published them)
                __syncwarp();
                const bool prefetched = tma_task && t0 >= kChunk && t0 + kChunk <= ncols;   // issued one chunk ago
                if (has_in && !in_ring) link_wait(&lin->produced, (uint32_t)(t0 + kChunk < ncols ? t0 + kChunk : ncols), &slot->abort);
#pragma unroll
                for (int u = 0; u < kChunk / 32; ++u) {
                    const int col = t0 + u * 32 + lane;
                    uint8_t c = (uint8_t)kPadCode;
                    uint2 bv = make_uint2(FLOOR0, EF0);
                    if (col < ncols) {
                        if (!prefetched)
                            c = p.codes[c_rev ? task.c_base - (uint64_t)col : task.c_base + (uint64_t)col];
                    }
                    if (has_in) {
                        // a ring entry is valid once it carries the parity of the lap this column belongs to
                        const uint32_t q = q_in + (uint32_t)col;
                        const uint2* const ptr = b_in + (q & m_in);
                        const uint32_t tag = ((q >> p.ring_shift) + 1u) & 1u;
                        bool stale = false;
                        if (col < ncols) {
                            bv = __ldcg(ptr);
                            stale = in_ring && ((bv.x >> 15) & 1u) != tag;
                        }
                        if (__any_sync(kFull, stale)) bv = ring_poll(ptr, tag, bv, stale, &slot->abort);
                        if (in_ring) bv.x &= 0xffff7fffu;
                    }
                    // lane 0 consumes entry `col` at step col as (H of row -1 for the next step's diagonal, F entering
                    // row 0): biases (col - 1 - tbase) * ext and (col - tbase) * ext
                    const uint32_t bb = (uint32_t)((col - tbase) * ext) * 65537u;
                    bv.x += bb - p.ext2; bv.y += bb;
                    if (!prefetched) {
                        rring[col & (kRing - 1)] = c;
                        rring[(col & (kRing - 1)) + kRing] = c;
                    }
                    bring[col & (kChunk - 1)] = bv;
                }
                if (has_in && in_ring) {          // the ring entries of this chunk may be overwritten now
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&lin->consumed) = (uint32_t)(t0 + kChunk);
                }
                if (prefetched) {               // the bulk copies issued one chunk ago have landed?
                    // votes make the outcome warp-uniform in the compiler's eyes
                    bool landed = __all_sync(kFull, mbar_wait_hint(mbar, tma_parity, 1000000u));
                    if (!landed) landed = __all_sync(kFull, mbar_wait_hint(mbar, tma_parity, 10000000u));
                    if (!landed) landed = __all_sync(kFull, mbar_wait_hint(mbar, tma_parity, 100000000u));
                    if (!landed) __trap();       // 256 bytes not delivered after > 100 ms: the device is broken
                    tma_parity ^= 1u;
                }
                __syncwarp();
                // ---- abandoned task (some tile saw a score >= hi)? No bulk copy is in flight at this point.
                if (hi) {
                    if (__any_sync(kFull, ld_vol(&slot->abort) != 0u)) { left_at = t0; cols_done = t0 < ncols ? t0 : ncols; break; }
                }
                // ---- TMA prefetch of the NEXT chunk of the read into the free ring slot (twice: the ring is mirrored);
                //      it lands while this chunk is computed. Needs 16-byte aligned addresses and a whole chunk; the
                //      begin pass (reversed read) and ragged tails use the plain loads above.
                if (tma_task && t0 + 2 * kChunk <= ncols) {
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order earlier LDS of that slot
                        const uint8_t* src = p.codes + task.c_base + (uint64_t)(t0 + kChunk);
                        uint8_t* dst = rring + ((t0 + kChunk) & (kRing - 1));
                        mbar_expect_tx(mbar, 2 * kChunk);
                        tma_load_1d(dst, src, kChunk, mbar);
                        tma_load_1d(dst + kRing, src, kChunk, mbar);
                    }
                }
                __syncwarp();

                const int t1 = (t0 + kChunk < total_steps) ? t0 + kChunk : total_steps;
                for (int g0 = t0; g0 < t1; g0 += 32) {
                    const int g1 = (g0 + 32 < t1) ? g0 + 32 : t1;     // g1 - g0 is even (total_steps is even)
                    for (int t = g0; t < g1; t += kUnroll) {        // g0, g1 and total_steps are multiples of kUnroll
                        uint32_t FL[R + kUnroll - 1];      // floors of (step t+u, row k) = FL[u + k]: warp-uniform sliding window
                        FL[0] = flb + (uint32_t)t * p.ext2;
#pragma unroll
                        for (int i = 1; i < R + kUnroll - 1; ++i)      // IMAD chain, FMA pipe
                            asm("mad.lo.u32 %0, %1, 65537, %2;" : "=r"(FL[i]) : "r"(ext_v), "r"(FL[i - 1]));
                        const char* bq = bsrc + (uint32_t)(t & (kChunk - 1)) * bstride;
                        const uint8_t* rq = rlane + (t & (kRing - 1));
                        uint2* sq = bstage + (t & 31);
#pragma unroll
                        for (int u = 0; u < kUnroll; u += 2) {      // fully unrolled: u is a constant in every copy
                            column(HA, HB, FL, bq, rq, sq, u);
                            column(HB, HA, FL, bq, rq, sq, u + 1);
                            track(HB, HA, t + u);
                        }
                    }
                    // ---- event filter: only a cell that reaches the final maximum of the whole alignment matters, and
                    //      any H seen so far is a lower bound of it. Lift every lane's running maximum to (warp-wide
                    //      maximum - 1): lanes then take the slow path only for values that can still win or tie.
                    {
                        uint32_t wb = bestprev;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            const uint32_t o = __shfl_xor_sync(kFull, wb, off);
                            wb = __vimax3_s16x2(wb, o, o);
                        }
                        // true maxima so far (the running maximum sits at the row-0 bias of step g1-1; no borrow: it is
                        // never below the floor). A parked half (unused B of a begin task) reads 255.
                        const uint32_t flg = flb + (uint32_t)(g1 - 1) * p.ext2;
                        const uint32_t tm = wb - flg;
                        sat |= ((tm & 0xffffu) > sat_thr ? 1u : 0u) | ((tm >> 16) > sat_thr ? 2u : 0u);
                        if (hi && lane == 0 && ((tm & 0xffffu) >= hi || (tm >> 16) >= hi))
                            *reinterpret_cast<volatile uint32_t*>(&slot->abort) = 1u;
                        // the tiles of the task share what they know: a cell of ANY tile is a lower bound of the result
                        const uint32_t l0 = ld_vol(&slot->lb[0]), l1 = ld_vol(&slot->lb[1]);
                        if (lane == 0) {
                            if ((tm & 0xffffu) > l0) atomicMax(&slot->lb[0], tm & 0xffffu);
                            if ((tm >> 16) > l1) atomicMax(&slot->lb[1], tm >> 16);
                        }
                        const uint32_t lbp = (l0 | (l1 << 16)) + flg;        // both < 32768 by the cap: no carry
                        wb = __vimax3_s16x2(wb, lbp, lbp);
                        const uint32_t lift = wb - 0x00010001u;      // both halves >= bias >= 1: no borrow
                        bestprev = __vimax3_s16x2(bestprev, lift, lift);
                    }
                    // ---- flush the staged boundary row of this group: entry q was produced at step g0+q for column
                    //      g0+q-31; remove its bias and store 32 columns with one coalesced 256-byte write
                    if (has_out) {
                        __syncwarp();
                        const int tp = g0 + lane;
                        const int col = tp - 31;
                        // Columns [0, cprev) were stored by the previous groups, a whole group ago: the fence has nothing
                        // to wait for, and the count is published one group late instead of stalling on this group's stores.
                        const int cprev = g0 - 31 < ncols ? (g0 - 31 > 0 ? g0 - 31 : 0) : ncols;
                        const int cmax = g1 - 31 < ncols ? (g1 - 31 > 0 ? g1 - 31 : 0) : ncols;
                        if (out_ring) {
                            // a ring holds ring_cols columns: do not overwrite what the consumer has not copied yet
                            link_wait(&lout->consumed, (uint32_t)cmax > p.ring_cols ? (uint32_t)cmax - p.ring_cols : 0u, &slot->abort);
                        } else {
                            __threadfence();
                            __syncwarp();
                            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&lout->produced) = (uint32_t)cprev;
                        }
                        if (tp < g1 && col >= 0 && col < ncols) {
                            uint2 v = bstage[lane];
                            const uint32_t bb = (uint32_t)((tp - tbase + R - 1) * ext) * 65537u;   // row R-1 at step tp
                            v.x -= bb; v.y -= bb + p.ext2;
                            const uint32_t q = q_out + (uint32_t)col;
                            if (out_ring) v.x = (v.x & 0xffff7fffu) | ((((q >> p.ring_shift) + 1u) & 1u) << 15);   // lap parity
                            __stcg(b_out + (q & m_out), v);
                        }
                        __syncwarp();
                    }
                }
            }

            if (left_at >= 0) {                  // abandoned: columns [0, left_at - 31) of the last row are out, the link is owed the rest
                const int from = left_at - 31 > 0 ? (left_at - 31 < ncols ? left_at - 31 : ncols) : 0;
                if (has_out) link_abort_fill(p, b_out, m_out, out_ring, q_out, from, ncols, lout);
            } else if (has_out && !out_ring) {   // the last group's stores, then the final count
                __threadfence();
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&lout->produced) = (uint32_t)ncols;
            }
            // ---- tile epilogue: warp arg-best per half, merged into the task's result (lane 0, under the block lock)
            const uint32_t fl = flb + (uint32_t)(total_steps - 1) * p.ext2;     // floor of the last step
            // Every lane resolves its slots, whether it wrote them in this tile or not (a per-lane "valid" flag would
            // cost the compiler its convergence analysis of the loop above): a lane that holds the warp-wide maximum rose
            // to it in this tile, so its slot is current; any other lane loses the arg-best below whatever it reads. A
            // maximum that only came in through the shared bound belongs to another tile, which reports it with its cell.
            int s0 = (int)(bestprev & 0xffffu) - (int)(fl & 0xffffu), c0 = 0, r0 = 0;
            resolve_snap(0, c0, r0);
            if (s0 <= 0) { s0 = 0; c0 = 0x7fffffff; r0 = 0x7fffffff; }
            warp_argbest(s0, c0, r0);
            int s1 = (int)(bestprev >> 16) - (int)(fl >> 16), c1 = 0, r1 = 0;
            resolve_snap(1, c1, r1);
            if (s1 <= 0) { s1 = 0; c1 = 0x7fffffff; r1 = 0x7fffffff; }
            warp_argbest(s1, c1, r1);
            if (lane == 0) {
                const int ra = a_len - tile * 32 * R, rb = b_len - tile * 32 * R;
                const int rows_here = (ra < 0 ? 0 : (ra > 32 * R ? 32 * R : ra)) + (rb < 0 ? 0 : (rb > 32 * R ? 32 * R : rb));
                finish_job(ctl, p, code, s0, c0, r0, s1, c1, r1, sat, (unsigned long long)rows_here * (unsigned long long)cols_done);
            }
            __syncwarp();
        }
    }
    __syncthreads();          // every warp of the block has left the job loop
    if (threadIdx.x < kRings) p.ring_pos[blockIdx.x * kRings + threadIdx.x] = ctl->stream[threadIdx.x];
}

}  // namespace nmsv
