// nmsv_api.cu -- C ABI (include/nmsv.h) and the small planning / decision kernels around the wavefront kernel.
//
// Device pipeline of one validation batch (nmsv_plan_run; everything asynchronous on one stream):
//   plan_fwd_kernel      (read, slot) -> SwTask: template + its reverse complement against the whole read
//   sw_wavefront_kernel  per strip class R: scores and end cells of both strands          [the hot kernel]
//   decide_kernel        strand rule (count_sread_by_alignment.py:151-160), integer threshold / margin / mapq tests
//                        (:327-346, :379-389); emits begin-pass tasks only for reads that can still be supporting
//   sw_wavefront_kernel  per strip class R on the reversed, bounded windows: begin cells  (parasail src/ssw.c)
//   final_kernel         1-based 6-tuples (:153-162), final supporting decision, per-SV atomic counts, records
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/nmsv.h"
#include "nmsv_sw.cuh"
#include "nmsv_jump.cuh"

using namespace nmsv;

// ------------------------------------------------------------------------------------------------------------------
// strip classes
// ------------------------------------------------------------------------------------------------------------------
#ifdef NMSV_EXP_R32
static const int kClassR[] = {4, 8, 13, 16, 19, 32};
#else
static const int kClassR[] = {4, 8, 13, 16, 19};
#endif
static const int kNumClasses = (int)(sizeof(kClassR) / sizeof(kClassR[0]));

static int choose_class(uint32_t len)
{
    // Cheapest strip height for a template of `len` rows. One warp-step of the kernel costs about 8.25 * R issue cycles
    // of DPX work (66 DPX per 16 rows, 2 cycles each) plus ~34 cycles that do not depend on R (shuffles, profile and
    // ring loads, floor chain, loop), and a task runs tiles(len, R) * (columns + 31) steps: padded rows are paid for,
    // but so are short strips. The tiles of a task run side by side on the 4 warps of a block: beyond 4 tiles a task
    // is dealt out in groups of 4 (every fourth boundary row goes through a whole-read buffer), so a tile count that is
    // not a multiple of 4 leaves part of a group to other tasks' phase, and strips taller than 16 rows run at 12
    // instead of 16 warps per SM -- measured on 6000-row templates: R = 16 (12 tiles) 6.54 TCUPS, R = 19 (10 tiles)
    // 5.77, R = 13 (15 tiles) 5.85. Ties -> taller strips.
    static const int forced = [] { const char* e = getenv("NMSV_FORCE_R"); return e ? atoi(e) : 0; }();   // measurement knob
    if (forced)
        for (int c = 0; c < kNumClasses; ++c) if (kClassR[c] == forced) return c;
    int best = -1;
    double best_cost = 0.0;
    for (int c = 0; c < kNumClasses; ++c) {
        const uint64_t tile = 32ull * kClassR[c];
        uint64_t tiles = (len + tile - 1) / tile;
        if (tiles > (uint64_t)kW) tiles = (tiles + kW - 1) / kW * kW;
        const double cost = (double)tiles * (8.25 * kClassR[c] + 34.0) * (kClassR[c] > 16 ? 1.10 : 1.0);
        if (best < 0 || cost <= best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

static uint32_t tiles_of(uint32_t len, int cls)
{
    uint32_t tile = 32u * kClassR[cls];
    return (len + tile - 1) / tile;
}

// ------------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct nmsv_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop;
    int sm_clock_khz = 0;
    int blocks_per_sm[8] = {0};
    void* snap = nullptr;      // per resident warp: column snapshots of the running-maximum tracker (nmsv_sw.cuh)
    void* rings = nullptr;     // per resident block: kRings boundary rings of ring_cols columns (nmsv_sw.cuh)
    void* ring_pos = nullptr;  // per ring: its stream position, carried from launch to launch (entries are lap-tagged)
    uint32_t ring_cols = 1024;
    int prune = 1;             // nmsv_set_option("prune"): skip alignments that cannot change any output
    int jump_bps = 0;          // resident blocks per SM of the sw_jump kernel (set once)
    void* jump_stage = nullptr;      // pinned staging buffer of nmsv_sw_jump_batch (traceback operations on their way back)
    size_t jump_stage_bytes = 0;
    mutable std::string error;
};

static int set_err(const nmsv_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return set_err((ctx), NMSV_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                         \
    } while (0)

template <int R>
static cudaError_t setup_class(int* blocks_per_sm)
{
    const size_t smem = block_smem_bytes<R>();
    cudaError_t e = cudaFuncSetAttribute(sw_wavefront_kernel<R, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sw_wavefront_kernel<R, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int a = 0, b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sw_wavefront_kernel<R, 3, 1>, kWarpsPerBlock * 32, smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sw_wavefront_kernel<R, -1, -1>, kWarpsPerBlock * 32, smem);
    *blocks_per_sm = a < b ? a : b;      // one grid size (and boundary scratch size) per class
    return e;
}

static cudaError_t setup_class_idx(int cls, int* bps)
{
    switch (kClassR[cls]) {
    case 4: return setup_class<4>(bps);
    case 8: return setup_class<8>(bps);
    case 13: return setup_class<13>(bps);
    case 16: return setup_class<16>(bps);
    case 19: return setup_class<19>(bps);
#ifdef NMSV_EXP_R32
    case 32: return setup_class<32>(bps);
#endif
    }
    return cudaErrorInvalidValue;
}

template <int R>
static void launch_class(const SwParams& p, int grid, cudaStream_t s)
{
    const size_t smem = block_smem_bytes<R>();
    if (p.open == 3 && p.ext == 1)      // the reference's gap penalties: instantiation with immediate operands
        sw_wavefront_kernel<R, 3, 1><<<grid, kWarpsPerBlock * 32, smem, s>>>(p);
    else
        sw_wavefront_kernel<R, -1, -1><<<grid, kWarpsPerBlock * 32, smem, s>>>(p);
}

static void launch_class_idx(int cls, const SwParams& p, int grid, cudaStream_t s)
{
    switch (kClassR[cls]) {
    case 4: launch_class<4>(p, grid, s); break;
    case 8: launch_class<8>(p, grid, s); break;
    case 13: launch_class<13>(p, grid, s); break;
    case 16: launch_class<16>(p, grid, s); break;
    case 19: launch_class<19>(p, grid, s); break;
#ifdef NMSV_EXP_R32
    case 32: launch_class<32>(p, grid, s); break;
#endif
    }
}

static void fill_packed(SwParams& sp)
{
    auto pack = [](int v) { return ((uint32_t)(v & 0xffff) << 16) | (uint32_t)(v & 0xffff); };
    sp.neg_open2 = pack(-sp.open);
    sp.ext_m_open2 = pack(sp.ext - sp.open);
    sp.ext2 = (uint32_t)sp.ext * 65537u;
}

extern "C" int nmsv_abi_version(void) { return NMSV_ABI_VERSION; }

extern "C" int nmsv_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" const char* nmsv_last_error(const nmsv_ctx* ctx)
{
    return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

extern "C" int nmsv_create(nmsv_ctx** out, int device)
{
    if (!out) return set_err(nullptr, NMSV_E_INVALID, "nmsv_create: ctx pointer is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return set_err(nullptr, NMSV_E_CUDA, "nmsv_create: no usable CUDA device (%s); there is no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return set_err(nullptr, NMSV_E_INVALID, "nmsv_create: device %d out of range [0,%d)", device, n);
    nmsv_ctx* ctx = new (std::nothrow) nmsv_ctx();
    if (!ctx) return set_err(nullptr, NMSV_E_INVALID, "nmsv_create: out of host memory");
    ctx->device = device;
    { const char* e = getenv("NMSV_PRUNE"); if (e && e[0] == '0') ctx->prune = 0; }      // measurement knob
#define CREATE_TRY(call)                                                                                       \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            set_err(nullptr, NMSV_E_CUDA, "nmsv_create: %s failed: %s", #call, cudaGetErrorString(e__));       \
            delete ctx;                                                                                        \
            return NMSV_E_CUDA;                                                                                \
        }                                                                                                      \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major < 9) {
        set_err(nullptr, NMSV_E_CUDA, "nmsv_create: device %s (sm_%d%d) has no DPX instructions; this library is built for sm_100a",
                ctx->prop.name, ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return NMSV_E_CUDA;
    }
    CREATE_TRY(cudaDeviceGetAttribute(&ctx->sm_clock_khz, cudaDevAttrClockRate, device));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (int c = 0; c < kNumClasses; ++c) CREATE_TRY(setup_class_idx(c, &ctx->blocks_per_sm[c]));
    // block-per-problem sw_jump kernel: three diagonals of the longest region it accepts (set once, not per call)
    CREATE_TRY(cudaFuncSetAttribute(sw_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(3ull * (kJumpMaxRegion + 2) * sizeof(int32_t))));
    {
        int g = 0;
        for (int c = 0; c < kNumClasses; ++c) g = std::max(g, ctx->prop.multiProcessorCount * std::max(1, ctx->blocks_per_sm[c]));
        CREATE_TRY(cudaMalloc(&ctx->snap, (size_t)g * kWarpsPerBlock * kSnapWords * 32 * sizeof(uint4)));
        // boundary rings between the tiles of a task: 1024 columns x 8 B x kRings per block = 24 MB for 592 blocks, far
        // below the L2 size, rewritten every few hundred microseconds (NMSV_RING_COLS: measurement knob, power of two)
        if (const char* e = getenv("NMSV_RING_COLS")) {
            const int v = atoi(e);
            if (v >= 256 && (v & (v - 1)) == 0) ctx->ring_cols = (uint32_t)v;
        }
        CREATE_TRY(cudaMalloc(&ctx->rings, (size_t)g * kRings * ctx->ring_cols * sizeof(uint2)));
        CREATE_TRY(cudaMalloc(&ctx->ring_pos, (size_t)g * kRings * sizeof(uint32_t)));
        // lap 0 writes parity 1: zeroed memory reads as "not written yet"
        CREATE_TRY(cudaMemset(ctx->rings, 0, (size_t)g * kRings * ctx->ring_cols * sizeof(uint2)));
        CREATE_TRY(cudaMemset(ctx->ring_pos, 0, (size_t)g * kRings * sizeof(uint32_t)));
    }
    {   // keep freed blocks cached in the stream-ordered pool: plans are created and destroyed per batch
        cudaMemPool_t pool;
        CREATE_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = ~0ull;
        CREATE_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
#undef CREATE_TRY
    *out = ctx;
    return NMSV_OK;
}

extern "C" void nmsv_destroy(nmsv_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->own_stream); cudaStreamDestroy(ctx->own_stream); }
    if (ctx->snap) cudaFree(ctx->snap);
    if (ctx->rings) cudaFree(ctx->rings);
    if (ctx->ring_pos) cudaFree(ctx->ring_pos);
    if (ctx->jump_stage) cudaFreeHost(ctx->jump_stage);
    delete ctx;
}

extern "C" int nmsv_set_option(nmsv_ctx* ctx, const char* name, int value)
{
    if (!ctx || !name) return NMSV_E_INVALID;
    if (!strcmp(name, "prune")) { ctx->prune = value != 0; return NMSV_OK; }
    return set_err(ctx, NMSV_E_INVALID, "nmsv_set_option: unknown option '%s'", name);
}

extern "C" int nmsv_set_stream(nmsv_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return NMSV_E_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NMSV_OK;
}

extern "C" int nmsv_device_info(const nmsv_ctx* ctx, int32_t* sm_count, int32_t* sm_clock_khz, int64_t* l2_bytes,
                                int64_t* global_mem_bytes, char* name, size_t name_len)
{
    if (!ctx) return NMSV_E_INVALID;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (sm_clock_khz) *sm_clock_khz = ctx->sm_clock_khz;
    if (l2_bytes) *l2_bytes = ctx->prop.l2CacheSize;
    if (global_mem_bytes) *global_mem_bytes = (int64_t)ctx->prop.totalGlobalMem;
    if (name && name_len) { strncpy(name, ctx->prop.name, name_len - 1); name[name_len - 1] = 0; }
    return NMSV_OK;
}

static const uint8_t* code_table()
{
    static uint8_t tab[256];
    static const bool init = [] {
        for (int i = 0; i < 256; ++i) tab[i] = 4;
        tab['A'] = tab['a'] = 0; tab['C'] = tab['c'] = 1; tab['G'] = tab['g'] = 2; tab['T'] = tab['t'] = 3;
        return true;
    }();
    (void)init;
    return tab;
}

// Host-side packing for callers that hold the reads as ASCII inside one large buffer (a seam TSV read or mapped in
// blocks): sequence i = src[src_off[i] .. +lens[i]) is encoded straight to dst[dst_off[i] ..), one pass, no
// intermediate copies. Bytes of dst between the sequences are left as they are (the caller pre-fills the padding).
static void encode_scalar(const uint8_t* in, uint8_t* out, uint64_t n)
{
    const uint8_t* tab = code_table();
    for (uint64_t i = 0; i < n; ++i) out[i] = tab[in[i]];
}

#if defined(__x86_64__)
// 32 bases per step: fold case (c | 0x20), four byte compares, code = 4 - 4*[a] - 3*[c] - 2*[g] - 1*[t]
__attribute__((target("avx2"))) static void encode_avx2(const uint8_t* in, uint8_t* out, uint64_t n)
{
    const __m256i lower = _mm256_set1_epi8(0x20), four = _mm256_set1_epi8(4);
    const __m256i ca = _mm256_set1_epi8('a'), cc = _mm256_set1_epi8('c'), cg = _mm256_set1_epi8('g'), ct = _mm256_set1_epi8('t');
    const __m256i k4 = _mm256_set1_epi8(4), k3 = _mm256_set1_epi8(3), k2 = _mm256_set1_epi8(2), k1 = _mm256_set1_epi8(1);
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_or_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i)), lower);
        __m256i r = four;
        r = _mm256_sub_epi8(r, _mm256_and_si256(_mm256_cmpeq_epi8(v, ca), k4));
        r = _mm256_sub_epi8(r, _mm256_and_si256(_mm256_cmpeq_epi8(v, cc), k3));
        r = _mm256_sub_epi8(r, _mm256_and_si256(_mm256_cmpeq_epi8(v, cg), k2));
        r = _mm256_sub_epi8(r, _mm256_and_si256(_mm256_cmpeq_epi8(v, ct), k1));
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i), r);
    }
    encode_scalar(in + i, out + i, n - i);
}
#endif

// (c | 0x20) also folds a few non-letters onto a/c/g/t: 'A'^0x20 pairs only -- 0x41/0x61, 0x43/0x63, 0x47/0x67, 0x54/0x74
// are the ONLY bytes whose (c | 0x20) equals a, c, g or t, so the vector path equals the table.
static void encode_bytes(const uint8_t* in, uint8_t* out, uint64_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { encode_avx2(in, out, n); return; }
#endif
    encode_scalar(in, out, n);
}

extern "C" void nmsv_encode_pack(const char* src, const uint64_t* src_off, const uint32_t* lens, const uint64_t* dst_off,
                                 uint64_t n_seqs, uint8_t* dst)
{
    for (uint64_t s = 0; s < n_seqs; ++s)
        encode_bytes(reinterpret_cast<const uint8_t*>(src) + src_off[s], dst + dst_off[s], lens[s]);
}

// the same, and every sequence is padded with the wildcard code up to the next 16-byte boundary (dst_off[] must be
// 16-byte aligned): the destination buffer needs no pre-fill
extern "C" void nmsv_encode_pack_padded(const char* src, const uint64_t* src_off, const uint32_t* lens,
                                        const uint64_t* dst_off, uint64_t n_seqs, uint8_t* dst)
{
    for (uint64_t s = 0; s < n_seqs; ++s) {
        uint8_t* out = dst + dst_off[s];
        const uint64_t n = lens[s];
        encode_bytes(reinterpret_cast<const uint8_t*>(src) + src_off[s], out, n);
        for (uint64_t i = n; i < ((n + 15) & ~15ull); ++i) out[i] = 4;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side of the stage seam: the group-by over the key-sorted seam TSV (count_sread_by_alignment.py:414-448)
// ------------------------------------------------------------------------------------------------------------------
static inline bool seam_ws(unsigned char c)      // bytes.strip() of the sequence / str.split() of the read id
{
    return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f);
}

static bool seam_mapq(const char* p, const char* e, int32_t* out)
{
    if (e - p == 4 && !memcmp(p, "None", 4)) { *out = NMSV_MAPQ_NONE; return true; }
    while (p < e && seam_ws((unsigned char)*p)) ++p;
    while (e > p && seam_ws((unsigned char)e[-1])) --e;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p == e) return false;
    long long v = 0;
    for (; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
        if (v > 0x7fffffffLL) return false;
    }
    *out = (int32_t)(neg ? -v : v);
    return true;
}

extern "C" int nmsv_seam_scan(const char* data, uint64_t size, uint64_t start, uint64_t max_bases, uint64_t tail_bytes, int is_sbnd,
                              nmsv_seam_group* groups, uint32_t group_cap, nmsv_seam_read* reads, uint32_t read_cap,
                              uint32_t* n_groups, uint32_t* n_reads, uint64_t* next, uint64_t* error_off)
{
    if (!data || !groups || !reads || !n_groups || !n_reads || !next) return NMSV_E_INVALID;
    uint32_t ng = 0, nr = 0;
    uint64_t bases = 0, pos = start;
    std::vector<int32_t> table;          // open addressing over the reads of the current group: index into reads, -1 = free
    auto id_hash = [&](const char* p, uint32_t n) {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
        return h;
    };
    auto rebuild = [&](uint32_t begin, uint32_t end, size_t cap) {
        table.assign(cap, -1);
        for (uint32_t r = begin; r < end; ++r) {
            size_t at = id_hash(data + reads[r].id_off, reads[r].id_len) & (cap - 1);
            while (table[at] >= 0) at = (at + 1) & (cap - 1);
            table[at] = (int32_t)r;
        }
    };
    *n_groups = 0; *n_reads = 0; *next = start;
    while (pos < size) {
        // ---- one key group
        const uint64_t group_start = pos;
        const uint32_t first_read = nr;
        const char* key = nullptr;
        uint32_t key_len = 0;
        uint64_t group_bases = 0;
        size_t cap = 64;
        table.assign(cap, -1);
        bool overflow = false;
        while (pos < size) {
            const char* line = data + pos;
            const char* le = (const char*)memchr(line, '\n', size - pos);
            if (!le) le = data + size;
            const char* t[4];
            const char* q = line;
            int nt = 0;
            while (nt < 4) {
                const char* tab = (const char*)memchr(q, '\t', le - q);
                if (!tab) break;
                t[nt++] = tab; q = tab + 1;
            }
            if (nt < 4 || memchr(q, '\t', le - q)) { if (error_off) *error_off = pos; return NMSV_E_INVALID; }
            const uint32_t klen = (uint32_t)(t[0] - line);
            if (!key) { key = line; key_len = klen; }
            else if (klen != key_len || memcmp(key, line, klen)) break;        // next group starts here
            // read id: first whitespace-separated token of the second field (the FASTA header round trip, pyssw.py:41)
            const char* ib = t[0] + 1;
            const char* ie = t[1];
            while (ib < ie && seam_ws((unsigned char)*ib)) ++ib;
            const char* ix = ib;
            while (ix < ie && !seam_ws((unsigned char)*ix)) ++ix;
            int32_t m1, m2 = NMSV_MAPQ_NONE;
            if (!seam_mapq(t[1] + 1, t[2], &m1) || (!is_sbnd && !seam_mapq(t[2] + 1, t[3], &m2))) {
                if (error_off) *error_off = pos;
                return NMSV_E_INVALID;
            }
            const char* sb = t[3] + 1;
            const char* se = le;
            while (se > sb && seam_ws((unsigned char)se[-1])) --se;
            while (sb < se && seam_ws((unsigned char)*sb)) ++sb;
            // the dict of the reference (:162, add_long_read_seq): the last record of an id gives its mapq; its sequence is
            // the last non-empty one; ids are ordered by their first non-empty record
            const uint32_t idl = (uint32_t)(ix - ib);
            size_t at = id_hash(ib, idl) & (cap - 1);
            int32_t hit = -1;
            while (table[at] >= 0) {
                const nmsv_seam_read& o = reads[table[at]];
                if (o.id_len == idl && !memcmp(data + o.id_off, ib, idl)) { hit = table[at]; break; }
                at = (at + 1) & (cap - 1);
            }
            if (hit >= 0) {
                nmsv_seam_read& o = reads[hit];
                o.mapq1 = m1; o.mapq2 = m2;
                if (se > sb) {
                    group_bases += (uint64_t)(se - sb); group_bases -= o.seq_len;
                    o.seq_off = (uint64_t)(sb - data); o.seq_len = (uint32_t)(se - sb);
                }
            } else if (se > sb) {
                if (nr >= read_cap) { overflow = true; break; }
                nmsv_seam_read& o = reads[nr];
                o.id_off = (uint64_t)(ib - data); o.id_len = idl;
                o.seq_off = (uint64_t)(sb - data); o.seq_len = (uint32_t)(se - sb);
                o.mapq1 = m1; o.mapq2 = m2;
                table[at] = (int32_t)nr;
                ++nr;
                group_bases += o.seq_len;
                if ((size_t)(nr - first_read) * 2 > cap) { cap *= 2; rebuild(first_read, nr, cap); }
            }
            pos = (uint64_t)(le - data) + 1;
        }
        if (overflow || ng >= group_cap) {       // this group does not fit: it belongs to the next call
            if (ng == 0) return NMSV_E_CAPACITY;
            nr = first_read;
            pos = group_start;
            break;
        }
        nmsv_seam_group& g = groups[ng++];
        g.key_off = (uint64_t)(key - data); g.key_len = key_len;
        g.read_begin = first_read; g.read_end = nr; g.reserved = 0;
        bases += group_bases;
        if (pos > size) pos = size;
        if (bases >= max_bases && size - std::min(pos, size) >= tail_bytes) break;      // a short rest joins this batch
    }
    *n_groups = ng; *n_reads = nr; *next = pos > size ? size : pos;
    return NMSV_OK;
}

extern "C" void nmsv_encode_ascii(const char* ascii, uint64_t n, uint8_t* codes)
{
    encode_bytes(reinterpret_cast<const uint8_t*>(ascii), codes, n);
}

// ------------------------------------------------------------------------------------------------------------------
// integer roofline: sustained DPX issue rate, measured on the device that runs the kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dpx_peak_kernel(uint32_t* out, uint32_t seed, uint32_t k1, int iters)
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 3u + i; b[i] = seed ^ (i * 77u); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __viaddmax_s16x2(a[i], k1, b[i]);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

extern "C" int nmsv_measure_dpx_peak(nmsv_ctx* ctx, double* lane_inst_per_s)
{
    if (!ctx || !lane_inst_per_s) return NMSV_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256, iters = 8192;
    uint32_t* d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(uint32_t) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        dpx_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 12345u, 0xfffdfffdu, iters);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    CUDA_TRY(ctx, cudaGetLastError());
    *lane_inst_per_s = (double)blocks * threads * iters * 8.0 / (best * 1e-3);
    return NMSV_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// device buffers (stream-ordered pool)
// ------------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t n, cudaStream_t s)
    {
        release();
        stream = s;
        bytes = n ? n : 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, s);
        if (e != cudaSuccess) { p = nullptr; bytes = 0; }
        return e;
    }
    void release()
    {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr; bytes = 0;
    }
    ~DevBuf() { release(); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ------------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------------
struct Sel { int32_t score, erow, ecol, strand; };   // chosen strand of one (read, slot)

// per strip class (index = position in the plan's class list, at most 8): forward queue head, begin queue head and
// begin queue reservation count
enum : uint32_t { kCntFwdQueue = 0, kCntRevQueue = 8, kCntRecords = 17, kCntStatus = 18, kCntSaturated = 19,
                  kCntRevReserve = 24, kCntTotal = 32 };
// device work accounting (unsigned long long each)
enum : uint32_t { kCellsFwd = 0, kCellsBegin = 1, kTasksSpawnedFwd = 2, kCellsTotal = 4 };
enum : uint8_t { kReadCandidate = 1, kReadSupporting = 2,
                 kReadPruned = 4,      // mapq test already failed: counted as checked, never aligned (:344-346, :388-389)
                 kReadVarDone = 8,     // lazy reference alleles: the variant tests passed, the first reference allele is queued
                 kReadRef1Done = 16,   // ... the margin test against it passed too, the second one is queued
                 kReadVarRedo = 32 };  // a variant segment that stayed below its threshold was re-aligned exactly (the read
                                       // passed through the other segment, so both tuples go into its record)

__device__ __forceinline__ int alias_of(const nmsv_sv& sv, int s)
{
    for (int q = 0; q < s; ++q)
        if (sv.tmpl[q] == sv.tmpl[s] && sv.tmpl_rc[q] == sv.tmpl_rc[s]) return q;
    return s;
}

__device__ __forceinline__ bool slot_present(const nmsv_sv& sv, int s)
{
    if (sv.tmpl[s] == NMSV_NONE) return false;
    if ((sv.flags & NMSV_SV_SBND) && (s == NMSV_SLOT_VAR2 || s == NMSV_SLOT_REF2)) return false;
    return true;
}

// The mapping-quality test of the decision rules (count_sread_by_alignment.py:344-346, sbnd :388-389). It does not
// depend on any alignment: a read that fails it is counted in `checked` and can never be supporting.
__device__ __forceinline__ bool mapq_ok(const nmsv_sv& sv, const nmsv_read& rd)
{
    if (sv.flags & NMSV_SV_SBND) return rd.mapq1 != NMSV_MAPQ_NONE && rd.mapq1 >= sv.min_mapq;
    return rd.mapq1 != NMSV_MAPQ_NONE && rd.mapq1 >= sv.min_mapq && rd.mapq2 != NMSV_MAPQ_NONE && rd.mapq2 >= sv.min_mapq;
}

// Reference-allele slots are only read by the margin tests (:336-341, :384-385), which only matter for a read that
// already passed the variant tests: with lazy_refs their alignments are queued by the continuation of the variant tasks.
__device__ __forceinline__ bool is_lazy_slot(const nmsv_sv& sv, int s)
{
    return (s == NMSV_SLOT_REF1 || s == NMSV_SLOT_REF2) && alias_of(sv, s) >= NMSV_SLOT_REF1;
}

// Score below which the alignment of variant slot s cannot matter unless another variant segment passes: the score test
// of the decision rules (:328-333). A task shared by var1 and var2 (identical templates) takes the smaller of the two.
__device__ __forceinline__ uint32_t var_threshold(const nmsv_sv& sv, int s)
{
    if (s > NMSV_SLOT_VAR2) return 0u;
    int t = sv.score_min[s];
    const int o = 1 - s;
    if (!(sv.flags & NMSV_SV_SBND) && sv.tmpl[o] == sv.tmpl[s] && sv.tmpl_rc[o] == sv.tmpl_rc[s] && sv.score_min[o] < t) t = sv.score_min[o];
    return t > 1 ? (uint32_t)t : 0u;
}

__device__ __forceinline__ SwTask make_fwd_task(const nmsv_sv& sv, int s, uint32_t out, uint32_t read_seq,
                                                const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ lens,
                                                const uint8_t* __restrict__ seq_class, uint32_t thr = 0u, uint32_t hi = 0u)
{
    SwTask t;
    memset(&t, 0, sizeof(t));
    const uint32_t tm = sv.tmpl[s];
    const uint32_t m = lens[tm];
    t.out = out;
    t.a_base = offsets[tm];
    t.a_len = m;
    const uint32_t rc = sv.tmpl_rc[s];
    if (rc == NMSV_NONE) { t.b_base = offsets[tm] + m - 1; t.flags = kBRev | kBComp; }
    else { t.b_base = offsets[rc]; }
    t.b_len = m;
    t.c_base = offsets[read_seq];
    t.ncols = lens[read_seq];
    t.rclass = (uint32_t)seq_class[tm];   // strip height R itself
    t.thr = thr;
    t.hi = hi;
    return t;
}

// thread per (rank, slot): rank = position of the read in the cost-sorted order
__global__ void plan_fwd_kernel(const nmsv_sv* __restrict__ svs, const nmsv_read* __restrict__ reads,
                                const uint32_t* __restrict__ order, const uint64_t* __restrict__ offsets,
                                const uint32_t* __restrict__ lens, const uint8_t* __restrict__ seq_class,
                                uint32_t n_reads, SwTask* __restrict__ tasks, uint32_t* __restrict__ group_remaining,
                                uint8_t* __restrict__ read_flags, int prune, int lazy_refs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads * 4u) return;
    const uint32_t r = order[i >> 2];
    const int s = (int)(i & 3u);
    const nmsv_read rd = reads[r];
    const nmsv_sv sv = svs[rd.sv];
    SwTask t;
    memset(&t, 0, sizeof(t));
    t.out = r * 4u + (uint32_t)s;
    t.rclass = 0xffffffffu;
    const bool pruned = prune && !mapq_ok(sv, rd);
    if (pruned && s == 0) read_flags[r] = kReadPruned;       // zeroed before the launch
    if (!pruned && slot_present(sv, s) && alias_of(sv, s) == s && !(lazy_refs && is_lazy_slot(sv, s))) {
        // with the continuation chain available, variant alignments are speculative: exact only from the score test up
        t = make_fwd_task(sv, s, r * 4u + (uint32_t)s, rd.seq, offsets, lens, seq_class, lazy_refs ? var_threshold(sv, s) : 0u);
        if (t.ncols) atomicAdd(&group_remaining[r], 1u);      // zeroed before the launch
    }
    tasks[i] = t;
}

struct DecideParams {
    const nmsv_sv* svs; const nmsv_read* reads; const uint64_t* offsets; const uint32_t* lens;
    const uint8_t* seq_class; uint32_t n_reads;
    const SwResult* fwd; Sel* sel; uint8_t* read_flags;
    // dynamic queues: one region of rev_capacity entries per strip class in use (region_of_r: R -> region index);
    // begin-pass tasks and, with lazy_refs, the forward tasks of the reference alleles
    SwTask* rev_tasks; uint32_t* rev_ready; uint32_t rev_capacity; uint32_t* counters;
    uint32_t* group_remaining; unsigned long long* cells;
    uint8_t region_of_r[kMaxR + 1];
    int match, open, ext;
    int lazy_refs;
};

namespace nmsv { struct GroupHook { DecideParams d; }; }

// rows / columns an alignment of `score` that ends at (erow, ecol) can span at most: every gap position costs at least
// min(open, ext), and an alignment over c columns (r rows) scores at most match * c (match * r)
// (DESIGN.md "begin pass window")
__device__ __forceinline__ uint32_t window_extent(int own, int other, int score, int match, int open, int ext)
{
    // own = extent available in this dimension (end coordinate + 1), other = extent available in the other one
    const int cmin = open < ext ? open : ext;
    long long w = (long long)own;
    if (cmin > 0) {
        long long slack = ((long long)match * other - score) / cmin;
        if (slack < 0) slack = 0;
        const long long bound = (long long)other + slack;
        if (bound < w) w = bound;
    }
    return (uint32_t)w;
}

// publish a task in the dynamic queue of its strip class: body, fence, ready flag (a launch of that class may be
// draining the queue right now)
__device__ __forceinline__ void push_task(const DecideParams& p, const SwTask& t)
{
    const uint32_t region = p.region_of_r[t.rclass];
    const uint32_t idx = atomicAdd(&p.counters[kCntRevReserve + region], 1u);
    const size_t at = (size_t)region * p.rev_capacity + idx;
    p.rev_tasks[at] = t;
    __threadfence();
    atomicExch(&p.rev_ready[at], 1u);
}

__device__ __forceinline__ void emit_rev_task(const DecideParams& p, uint32_t out, uint32_t tm, uint32_t rc,
                                              uint32_t read_seq, const Sel& sl)
{
    const uint32_t m = p.lens[tm];
    SwTask t;
    memset(&t, 0, sizeof(t));
    // Every optimal alignment that ends in the end cell lies inside the window [erow - rows + 1, erow] x
    // [ecol - cols + 1, ecol]; cells of the reversed prefixes inside a rectangle anchored at the end cell have the
    // same values as in parasail's full second pass, cells outside cannot reach the score.
    const uint32_t rows = window_extent(sl.erow + 1, sl.ecol + 1, sl.score, p.match, p.open, p.ext);
    t.a_len = rows;
    if (sl.strand > 0) { t.a_base = p.offsets[tm] + (uint64_t)sl.erow; t.flags = kARev; }               // reverse(T[0..e])
    else if (rc == NMSV_NONE) { t.a_base = p.offsets[tm] + (uint64_t)(m - 1 - sl.erow); t.flags = kAComp; } // reverse(RC(T)[0..e])
    else { t.a_base = p.offsets[rc] + (uint64_t)sl.erow; t.flags = kARev; }
    t.b_len = 0;
    t.c_base = p.offsets[read_seq] + (uint64_t)sl.ecol;
    t.flags |= kCRev | kTaskBegin;
    t.ncols = window_extent(sl.ecol + 1, sl.erow + 1, sl.score, p.match, p.open, p.ext);
    t.out = out;
    t.best0 = (uint32_t)sl.score;       // only the cells that reach the forward score can be the begin cell
    t.rclass = (uint32_t)p.seq_class[tm];
    push_task(p, t);
}

__device__ __forceinline__ Sel select_strand(const SwResult* fwd_slot)
{
    const int2* fp = reinterpret_cast<const int2*>(fwd_slot);
    const int2 f0 = __ldcg(fp), f1 = __ldcg(fp + 1), f2 = __ldcg(fp + 2);      // results of other SMs: through L2
    Sel x;
    if (f0.x < 0 || f0.y < 0) { x.score = -1; x.erow = 0; x.ecol = 0; x.strand = 0; }   // saturated (parasail: None)
    else if (f0.x > f0.y) { x.score = f0.x; x.erow = f1.x; x.ecol = f2.x; x.strand = 1; }
    else { x.score = f0.y; x.erow = f1.y; x.ecol = f2.y; x.strand = -1; }                // ties -> '-'
    return x;
}

// The decision for one read once the forward alignments it is waiting for are known. Runs either as a thread of
// decide_kernel or, on the fused path, on lane 0 of the warp that finished the read's last outstanding forward task.
// With lazy_refs it runs up to three times for a read whose SV has reference-allele slots: after the variant alignments
// (variant tests; a read that passes gets its first reference-allele alignment queued), after that one (the margin
// test against it can already fail the read, :336-341: then the second reference allele is never aligned) and after the
// second one.
__device__ __noinline__ void decide_read(const DecideParams& p, const uint32_t r)
{
    const nmsv_read rd = p.reads[r];
    const nmsv_sv sv = p.svs[rd.sv];
    const uint8_t state = p.read_flags[r];        // written only by this read's own chain of continuations
    bool present[4];
    int lazy_slot[2] = {-1, -1};                  // reference-allele slots that are aligned lazily, in the order they are queued
    int n_lazy = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        present[s] = slot_present(sv, s);
        if (p.lazy_refs && present[s] && is_lazy_slot(sv, s) && alias_of(sv, s) == s) lazy_slot[n_lazy++] = s;
    }
    const int n_done = !p.lazy_refs ? 2 : ((state & kReadRef1Done) ? 2 : ((state & kReadVarDone) ? 1 : 0));   // lazy results in
    auto computed = [&](int s) {                  // is the alignment of slot s available now?
        if (!p.lazy_refs || !is_lazy_slot(sv, s)) return true;
        const int a = alias_of(sv, s);
        return (a == lazy_slot[0] && n_done >= 1) || (a == lazy_slot[1] && n_done >= 2);
    };
    Sel sl[4];
    bool have[4];
    bool saturated = false;
    int inexact = -1;                             // a variant slot with a task of its own whose result is only "below threshold"
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        Sel x = {0, 0, 0, 0};
        have[s] = present[s] && computed(s);
        if (have[s]) {
            const int a = alias_of(sv, s);
            x = select_strand(p.fwd + r * 4u + (uint32_t)a);
            saturated |= x.score < 0;
            if (p.lazy_refs && !(state & kReadVarRedo) && s <= NMSV_SLOT_VAR2 && x.score >= 0 &&
                (uint32_t)x.score < var_threshold(sv, a)) {
                x.strand = 0;                     // not an alignment: the speculative task only says "below the threshold"
                if (a == s) inexact = s;
            }
        }
        sl[s] = x;
        p.sel[r * 4u + s] = x;
    }
    if (saturated) { atomicAdd(&p.counters[kCntSaturated], 1u); p.read_flags[r] = 0; return; }
    const bool sbnd = sv.flags & NMSV_SV_SBND;
    bool pass = false;
    const int nvar = sbnd ? 1 : 2;
    const bool needs_margin = sbnd || (sv.flags & NMSV_SV_SHORT_DEL_DUP);
    // a lazily aligned reference allele was abandoned at (best variant score - margin + 1): such a result only says "within
    // the margin", it is not an alignment (the read fails the margin test below)
    if (p.lazy_refs && needs_margin) {
        int best_var = 0;
        for (int v = 0; v < nvar; ++v) if (present[v] && sl[v].score > best_var) best_var = sl[v].score;
        for (int s = NMSV_SLOT_REF1; s <= NMSV_SLOT_REF2; ++s)
            if (have[s] && is_lazy_slot(sv, s) && sl[s].score >= best_var - sv.margin + 1) { sl[s].strand = 0; p.sel[r * 4u + s].strand = 0; }
    }
    for (int v = 0; v < nvar; ++v) {
        if (!present[v] || sl[v].score < sv.score_min[v]) continue;
        const int m = (int)p.lens[sv.tmpl[v]];
        if (sl[v].strand > 0) pass |= (sl[v].erow + 1 >= sv.qend_min[v]);        // qend   = read_end1 + 1
        else pass |= (m - sl[v].erow <= sv.qstart_max[v]);                       // qstart = m - read_end1_r
    }
    pass = pass && mapq_ok(sv, rd);
    // margin test (:336-341, :384-385) against the reference alleles that are known so far: some variant segment must
    // beat every one of them by the margin. With all of them known this is the reference's test.
    if (pass && needs_margin) {
        bool any = false;
        for (int v = 0; v < nvar; ++v) {
            bool ok = present[v];
            for (int s = NMSV_SLOT_REF1; s <= NMSV_SLOT_REF2; ++s)
                if (have[s]) ok = ok && sl[v].score >= sl[s].score + sv.margin;
            any |= ok;
        }
        pass = any;
    }
    if (!pass) { p.read_flags[r] = 0; return; }
    if (inexact >= 0 && p.lens[rd.seq]) {
        // The read passes through the other variant segment, so the record will hold this segment's tuple too: align it
        // again, exactly. (Both segments below their thresholds never get here; a passing segment is exact.)
        p.read_flags[r] = (uint8_t)(state | kReadVarRedo);
        p.group_remaining[r] = 1u;
        atomicAdd(&p.cells[kTasksSpawnedFwd], 1ull);
        push_task(p, make_fwd_task(sv, inexact, r * 4u + (uint32_t)inexact, rd.seq, p.offsets, p.lens, p.seq_class));
        return;
    }
    const int n_avail = n_done < n_lazy ? n_done : n_lazy;
    if (p.lazy_refs && n_avail < n_lazy && p.lens[rd.seq]) {
        // still a candidate: its next reference-allele alignment is needed (for the margin test, or just for the record)
        const int s = lazy_slot[n_avail];
        // Where the margin test applies, all that matters about a reference allele is whether it scores within the
        // margin of the best variant segment: some segment beats it iff its score is below max_v(score_v) - margin + 1.
        // The task is abandoned as soon as a cell reaches that value (the read fails, nothing of it is recorded).
        uint32_t hi = 0u;
        if (needs_margin) {
            int best_var = 0;
            for (int v = 0; v < nvar; ++v) if (present[v] && sl[v].score > best_var) best_var = sl[v].score;
            const int b = best_var - sv.margin + 1;
            if (b <= 0) { p.read_flags[r] = 0; return; }      // every reference-allele score fails the margin test
            hi = (uint32_t)b;
        }
        p.read_flags[r] = (uint8_t)((state & kReadVarRedo) | (n_avail == 0 ? kReadVarDone : (kReadVarDone | kReadRef1Done)));
        p.group_remaining[r] = 1u;             // published with the task (fence inside push_task)
        atomicAdd(&p.cells[kTasksSpawnedFwd], 1ull);
        push_task(p, make_fwd_task(sv, s, r * 4u + (uint32_t)s, rd.seq, p.offsets, p.lens, p.seq_class, 0u, hi));
        return;
    }
    if (sbnd) pass = present[NMSV_SLOT_REF1];
    else if (sv.flags & NMSV_SV_SHORT_DEL_DUP) pass = present[NMSV_SLOT_REF1] && present[NMSV_SLOT_REF2];
    p.read_flags[r] = pass ? kReadCandidate : 0;
    if (!pass) return;
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (present[s] && alias_of(sv, s) == s && sl[s].score > 0)
            emit_rev_task(p, r * 4u + (uint32_t)s, sv.tmpl[s], sv.tmpl_rc[s], rd.seq, sl[s]);
}

__global__ void decide_kernel(const DecideParams p)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p.n_reads && !(p.read_flags[r] & kReadPruned)) decide_read(p, r);
}

__device__ void nmsv::sw_group_done(const GroupHook* hook, uint32_t out) { decide_read(hook->d, out >> 2); }

__device__ __forceinline__ nmsv_aln make_aln(const Sel& sl, int m, int brow, int bcol)
{
    nmsv_aln a;
    a.score = sl.score;
    a.strand = sl.strand;
    if (sl.strand > 0) { a.qstart = brow + 1; a.qend = sl.erow + 1; }
    else { a.qstart = m - sl.erow; a.qend = m - brow; }
    a.tstart = bcol + 1;
    a.tend = sl.ecol + 1;
    return a;
}

struct FinalParams {
    const nmsv_sv* svs; const nmsv_read* reads; const uint32_t* lens; uint32_t n_reads;
    const Sel* sel; const SwResult* rev; uint8_t* read_flags; nmsv_aln* alns;
    int32_t* supporting; nmsv_record* records; uint32_t record_capacity; uint32_t* counters;
};

__global__ void final_kernel(const FinalParams p)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads) return;
    const nmsv_read rd = p.reads[r];
    const nmsv_sv sv = p.svs[rd.sv];
    const bool cand = p.read_flags[r] & kReadCandidate;
    nmsv_aln al[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        nmsv_aln a = {0, 0, 0, 0, 0, 0};
        const Sel sl = p.sel[r * 4u + s];
        if (slot_present(sv, s) && sl.strand != 0) {       // strand 0: never aligned (pruned read / lazy reference allele)
            const int m = (int)p.lens[sv.tmpl[s]];
            int brow = 0, bcol = 0;
            if (cand && sl.score > 0) {
                const SwResult rv = p.rev[r * 4u + (uint32_t)alias_of(sv, s)];
                if (rv.score[0] != sl.score || rv.col[0] == 0x7fffffff)                // window bound violated: no cell
                    atomicAdd(&p.counters[kCntStatus], 1u);                            // of the window reached the score
                brow = sl.erow - rv.row[0];
                bcol = sl.ecol - rv.col[0];
            }
            a = make_aln(sl, m, brow, bcol);
            if (!cand) {   // begin pass skipped: begin-derived fields are not available
                if (sl.strand > 0) a.qstart = 0; else a.qend = 0;
                a.tstart = 0;
            }
        }
        al[s] = a;
        p.alns[r * 4u + s] = a;
    }
    if (!cand) return;
    bool supp = false;
    const int nvar = (sv.flags & NMSV_SV_SBND) ? 1 : 2;
    for (int v = 0; v < nvar; ++v)
        if (al[v].strand != 0)
            supp |= (al[v].score >= sv.score_min[v] && al[v].qstart <= sv.qstart_max[v] && al[v].qend >= sv.qend_min[v]);
    if (!supp) return;
    p.read_flags[r] = kReadCandidate | kReadSupporting;
    atomicAdd(&p.supporting[rd.sv], 1);
    const uint32_t idx = atomicAdd(&p.counters[kCntRecords], 1u);
    if (idx < p.record_capacity) {
        nmsv_record rec;
        rec.read = r; rec.sv = rd.sv;
#pragma unroll
        for (int s = 0; s < 4; ++s) rec.aln[s] = al[s];
        p.records[idx] = rec;
    }
}

// ---- pair API helpers ---------------------------------------------------------------------------------------------
struct PairParams {
    const uint64_t* offsets; const uint32_t* lens; const uint8_t* seq_class;
    const uint32_t* tmpl_idx; const uint32_t* rc_idx; const uint32_t* read_idx; uint32_t n_pairs;
    int both_strands;
    const SwResult* fwd; Sel* sel; SwTask* rev_tasks; uint32_t* rev_ready; uint32_t rev_capacity;
    uint8_t region_of_r[kMaxR + 1];
    const SwResult* rev; uint32_t* counters;
    nmsv_ssw_result* out_ssw; nmsv_aln* out_aln; int want_begin;
    int match, open, ext;
};

__global__ void pair_plan_kernel(const PairParams p, SwTask* __restrict__ tasks)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_pairs) return;
    const uint32_t tm = p.tmpl_idx[i], rs = p.read_idx[i];
    const uint32_t m = p.lens[tm];
    SwTask t;
    memset(&t, 0, sizeof(t));
    t.a_base = p.offsets[tm];
    t.a_len = m;
    if (p.both_strands) {
        const uint32_t rc = p.rc_idx ? p.rc_idx[i] : NMSV_NONE;
        if (rc == NMSV_NONE) { t.b_base = p.offsets[tm] + m - 1; t.flags = kBRev | kBComp; }
        else t.b_base = p.offsets[rc];
        t.b_len = m;
    }
    t.c_base = p.offsets[rs];
    t.ncols = p.lens[rs];
    t.out = i;
    t.rclass = (uint32_t)p.seq_class[tm];
    tasks[i] = t;
}

__global__ void pair_decide_kernel(const PairParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_pairs) return;
    const SwResult f = p.fwd[i];
    Sel x;
    if (f.score[0] < 0 || (p.both_strands && f.score[1] < 0)) { x.score = -1; x.erow = 0; x.ecol = 0; x.strand = 0; }   // saturated
    else if (!p.both_strands || f.score[0] > f.score[1]) { x.score = f.score[0]; x.erow = f.row[0]; x.ecol = f.col[0]; x.strand = 1; }
    else { x.score = f.score[1]; x.erow = f.row[1]; x.ecol = f.col[1]; x.strand = -1; }
    p.sel[i] = x;
    if (!p.want_begin || x.score <= 0) return;
    DecideParams d;
    d.offsets = p.offsets; d.lens = p.lens; d.seq_class = p.seq_class; d.rev_tasks = p.rev_tasks; d.counters = p.counters;
    d.rev_ready = p.rev_ready; d.rev_capacity = p.rev_capacity;
    for (int q = 0; q <= kMaxR; ++q) d.region_of_r[q] = p.region_of_r[q];
    d.match = p.match; d.open = p.open; d.ext = p.ext;
    emit_rev_task(d, i, p.tmpl_idx[i], p.rc_idx ? p.rc_idx[i] : NMSV_NONE, p.read_idx[i], x);
}

__global__ void pair_final_kernel(const PairParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_pairs) return;
    const Sel x = p.sel[i];
    if (x.score < 0) {      // parasail.ssw returns None: the score is not representable in the 16-bit lanes
        if (p.out_ssw) p.out_ssw[i] = nmsv_ssw_result{-1, -1, -1, -1, -1};
        if (p.out_aln) p.out_aln[i] = nmsv_aln{-1, 0, 0, 0, 0, 0};
        return;
    }
    int brow = 0, bcol = 0;
    if (p.want_begin && x.score > 0) {
        const SwResult rv = p.rev[i];
        if (rv.score[0] != x.score || rv.col[0] == 0x7fffffff) atomicAdd(&p.counters[kCntStatus], 1u);
        brow = x.erow - rv.row[0];
        bcol = x.ecol - rv.col[0];
    }
    if (p.out_ssw) {
        nmsv_ssw_result o;
        o.score1 = x.score; o.read_end1 = x.erow; o.ref_end1 = x.ecol;
        o.read_begin1 = p.want_begin ? brow : -1;
        o.ref_begin1 = p.want_begin ? bcol : -1;
        p.out_ssw[i] = o;
    }
    if (p.out_aln) p.out_aln[i] = make_aln(x, (int)p.lens[p.tmpl_idx[i]], brow, bcol);
}

// ------------------------------------------------------------------------------------------------------------------
// shared host-side validation / upload of the sequence tables
// ------------------------------------------------------------------------------------------------------------------
struct SeqTables {
    DevBuf codes, offsets, lens, seq_class;
    std::vector<uint8_t> cls_host;      // strip height R per sequence (0 = never used as a template)
    uint64_t h2d = 0;
};

static int check_scoring(nmsv_ctx* ctx, const nmsv_scoring* sc)
{
    if (!sc) return set_err(ctx, NMSV_E_INVALID, "scoring is NULL");
    if (sc->match <= 0 || sc->mismatch > 0 || sc->gap_open < 0 || sc->gap_extend < 0)
        return set_err(ctx, NMSV_E_INVALID, "scoring must have match > 0, mismatch <= 0, gap_open >= 0, gap_extend >= 0");
    if (sc->gap_open + sc->gap_extend > 64 || -sc->mismatch > 64 || sc->match > 64)
        return set_err(ctx, NMSV_E_RANGE, "scoring parameters too large for 16-bit lanes");
    return NMSV_OK;
}

static int check_seq_tables(nmsv_ctx* ctx, const uint8_t* codes, uint64_t n_codes, const uint64_t* offsets,
                            const uint32_t* lens, uint32_t n_seqs)
{
    if (!codes || !offsets || !lens) return set_err(ctx, NMSV_E_INVALID, "codes/offsets/lens must not be NULL");
    for (uint32_t s = 0; s < n_seqs; ++s)
        if (offsets[s] != NMSV_SEQ_ABSENT && (offsets[s] > n_codes || (uint64_t)lens[s] > n_codes - offsets[s]))
            return set_err(ctx, NMSV_E_INVALID, "sequence %u (offset %llu, len %u) lies outside the %llu-byte code buffer",
                           s, (unsigned long long)offsets[s], lens[s], (unsigned long long)n_codes);
    return NMSV_OK;
}

static int check_template(nmsv_ctx* ctx, uint32_t idx, const uint64_t* offsets, const uint32_t* lens, uint32_t n_seqs, const nmsv_scoring* sc, const char* what)
{
    if (idx >= n_seqs) return set_err(ctx, NMSV_E_INVALID, "%s index %u out of range (n_seqs %u)", what, idx, n_seqs);
    if (lens[idx] == 0) return set_err(ctx, NMSV_E_INVALID, "%s %u is empty", what, idx);
    if (offsets[idx] == NMSV_SEQ_ABSENT) return set_err(ctx, NMSV_E_INVALID, "%s %u has no sequence (NMSV_SEQ_ABSENT)", what, idx);
    // No limit on the length: parasail returns NULL only when a score actually leaves the 16-bit range, and so does the
    // kernel (a task whose bound match * min(rows, columns) exceeds sw_score_cap is watched at run time, nmsv_sw.cuh).
    (void)sc;
    return NMSV_OK;
}

// The sequence bytes are by far the largest upload: it is issued first, so that the DMA runs while the host validates
// the SV / read tables and sorts the reads; the strip classes (a result of that host work) follow in a second call.
static int upload_seq_codes(nmsv_ctx* ctx, SeqTables& st, const uint8_t* codes, uint64_t n_codes,
                            const uint64_t* offsets, const uint32_t* lens, uint32_t n_seqs)
{
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, st.codes.alloc(n_codes, s));
    CUDA_TRY(ctx, st.offsets.alloc(sizeof(uint64_t) * n_seqs, s));
    CUDA_TRY(ctx, st.lens.alloc(sizeof(uint32_t) * n_seqs, s));
    if (n_codes) CUDA_TRY(ctx, cudaMemcpyAsync(st.codes.p, codes, n_codes, cudaMemcpyHostToDevice, s));
    if (n_seqs) {
        CUDA_TRY(ctx, cudaMemcpyAsync(st.offsets.p, offsets, sizeof(uint64_t) * n_seqs, cudaMemcpyHostToDevice, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(st.lens.p, lens, sizeof(uint32_t) * n_seqs, cudaMemcpyHostToDevice, s));
    }
    st.h2d = n_codes + (uint64_t)n_seqs * (8 + 4 + 1);
    return NMSV_OK;
}

static int upload_seq_classes(nmsv_ctx* ctx, SeqTables& st, uint32_t n_seqs)
{
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, st.seq_class.alloc(n_seqs, s));
    if (n_seqs) CUDA_TRY(ctx, cudaMemcpyAsync(st.seq_class.p, st.cls_host.data(), n_seqs, cudaMemcpyHostToDevice, s));
    return NMSV_OK;
}

static int grid_for(const nmsv_ctx* ctx, int cls)
{
    return ctx->prop.multiProcessorCount * std::max(1, ctx->blocks_per_sm[cls]);
}

// global boundary buffers: kGBufs per block, only needed by tasks with more tiles than a block has warps
static int max_grid_gbufs(const nmsv_ctx* ctx)
{
    int g = 0;
    for (int c = 0; c < kNumClasses; ++c) g = std::max(g, grid_for(ctx, c));
    return g * kGBufs;
}

// ------------------------------------------------------------------------------------------------------------------
// validation plan
// ------------------------------------------------------------------------------------------------------------------
struct nmsv_plan {
    nmsv_ctx* ctx = nullptr;
    uint32_t n_seqs = 0, n_svs = 0, n_reads = 0;
    nmsv_scoring sc{};
    SeqTables st;
    DevBuf svs, reads, order, tasks, fwd, rev_tasks, rev_ready, group_remaining, hook, rev, sel, alns, flags, supporting,
           records, counters, cells, bnd;
    bool fused_begin = true;      // begin tasks are drained by the forward launches (NMSV_FUSED_BEGIN=0: separate launches only)
    bool prune = true;            // reads whose mapq test fails get no alignments (nmsv_set_option "prune")
    bool lazy_refs = true;        // ref1/ref2 alignments only for reads that pass the variant tests (needs the fused path)
    uint32_t rev_capacity = 0;    // entries per strip class of the dynamic queue
    uint32_t bnd_stride = 1;
    std::vector<int> classes;          // indices into kClassR that occur in this batch
    std::vector<int32_t> checked;
    nmsv_stats stats{};
    bool ran = false;
    // per-launch CUDA-event timing (nmsv_plan_set_profiling): events are recorded on the launching stream; the last
    // kProfRuns runs keep their own events, so that the durations can be averaged over the runs of a timed region
    // without a host synchronisation between them
    static constexpr uint32_t kProfRuns = 32, kProfSlots = 40;
    bool profiling = false;
    uint32_t prof_runs = 0;            // runs recorded since profiling was switched on
    std::vector<cudaEvent_t> events;   // [kProfRuns][kProfSlots], created on demand
    std::vector<uint32_t> launch_phase, launch_rclass;
    ~nmsv_plan() { for (cudaEvent_t e : events) if (e) cudaEventDestroy(e); }
};

enum : uint32_t { kPhasePlan = 0, kPhaseFwd = 1, kPhaseDecide = 2, kPhaseRev = 3, kPhaseFinal = 4 };

static void mark(nmsv_plan* pl, size_t& slot, cudaStream_t s)
{
    if (!pl->profiling || slot >= nmsv_plan::kProfSlots) return;
    if (pl->events.empty()) pl->events.assign((size_t)nmsv_plan::kProfRuns * nmsv_plan::kProfSlots, nullptr);
    cudaEvent_t& e = pl->events[(size_t)(pl->prof_runs % nmsv_plan::kProfRuns) * nmsv_plan::kProfSlots + slot++];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, s);
}

extern "C" void nmsv_plan_destroy(nmsv_plan* plan)
{
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    delete plan;
}

extern "C" int nmsv_plan_create(nmsv_ctx* ctx, nmsv_plan** out, const uint8_t* codes, uint64_t n_codes,
                                const uint64_t* offsets, const uint32_t* lens, uint32_t n_seqs, const nmsv_sv* svs,
                                uint32_t n_svs, const nmsv_read* reads, uint32_t n_reads, const nmsv_scoring* scoring)
{
    if (!ctx) return NMSV_E_INVALID;
    if (!out) return set_err(ctx, NMSV_E_INVALID, "plan pointer is NULL");
    *out = nullptr;
    int rc = check_scoring(ctx, scoring);
    if (rc) return rc;
    rc = check_seq_tables(ctx, codes, n_codes, offsets, lens, n_seqs);
    if (rc) return rc;
    if ((n_svs && !svs) || (n_reads && !reads)) return set_err(ctx, NMSV_E_INVALID, "svs/reads must not be NULL");
    if (n_reads > 0x3fffffffu) return set_err(ctx, NMSV_E_INVALID, "too many reads in one batch (%u)", n_reads);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));

    nmsv_plan* pl = new (std::nothrow) nmsv_plan();
    if (!pl) return set_err(ctx, NMSV_E_INVALID, "out of host memory");
    pl->ctx = ctx; pl->n_seqs = n_seqs; pl->n_svs = n_svs; pl->n_reads = n_reads; pl->sc = *scoring;
    // on an error return the upload from the caller's buffers may still be in flight: wait for it before they go away
    struct Guard { nmsv_plan* p; ~Guard() { if (p) { cudaStreamSynchronize(p->ctx->stream); nmsv_plan_destroy(p); } } } guard{pl};

    rc = upload_seq_codes(ctx, pl->st, codes, n_codes, offsets, lens, n_seqs);      // DMA in flight during the host work below
    if (rc) return rc;

    // ---- validate SVs, pick strip classes
    pl->st.cls_host.assign(n_seqs, 0);
    bool class_used[8] = {false};
    pl->checked.assign(n_svs, 0);
    for (uint32_t i = 0; i < n_svs; ++i) {
        const nmsv_sv& sv = svs[i];
        if (sv.read_begin > sv.read_end || sv.read_end > n_reads)
            return set_err(ctx, NMSV_E_INVALID, "SV %u has an invalid read range [%u,%u)", i, sv.read_begin, sv.read_end);
        pl->checked[i] = (int32_t)(sv.read_end - sv.read_begin);
        for (int s = 0; s < 4; ++s) {
            if (sv.tmpl[s] == NMSV_NONE) continue;
            rc = check_template(ctx, sv.tmpl[s], offsets, lens, n_seqs, scoring, "template");
            if (rc) return rc;
            if (sv.tmpl_rc[s] != NMSV_NONE) {
                rc = check_template(ctx, sv.tmpl_rc[s], offsets, lens, n_seqs, scoring, "reverse-complement template");
                if (rc) return rc;
                if (lens[sv.tmpl_rc[s]] != lens[sv.tmpl[s]])
                    return set_err(ctx, NMSV_E_INVALID, "SV %u slot %d: explicit reverse complement has a different length", i, s);
            }
            const int c = choose_class(lens[sv.tmpl[s]]);
            pl->st.cls_host[sv.tmpl[s]] = (uint8_t)kClassR[c];
            class_used[c] = true;
        }
        if (sv.tmpl[NMSV_SLOT_VAR1] == NMSV_NONE)
            return set_err(ctx, NMSV_E_INVALID, "SV %u has no variant template (var1)", i);
        if ((sv.flags & NMSV_SV_SBND) && sv.tmpl[NMSV_SLOT_REF1] == NMSV_NONE)
            return set_err(ctx, NMSV_E_INVALID, "single-breakend SV %u needs var1 and ref1", i);
        if ((sv.flags & NMSV_SV_SHORT_DEL_DUP) && !(sv.flags & NMSV_SV_SBND) &&
            (sv.tmpl[NMSV_SLOT_REF1] == NMSV_NONE || sv.tmpl[NMSV_SLOT_REF2] == NMSV_NONE))
            return set_err(ctx, NMSV_E_INVALID, "short del/dup SV %u needs ref1 and ref2", i);
    }
    for (int c = 0; c < kNumClasses; ++c) if (class_used[c]) pl->classes.push_back(c);

    {
        const char* e = getenv("NMSV_FUSED_BEGIN");
        pl->fused_begin = !(e && e[0] == '0');
    }
    pl->prune = ctx->prune != 0;
    pl->lazy_refs = pl->prune && pl->fused_begin;

    // ---- validate reads, cost-sort, work accounting
    std::vector<uint64_t> cost(n_reads);
    uint32_t max_cols_multi = 0;
    nmsv_stats& stt = pl->stats;
    for (uint32_t r = 0; r < n_reads; ++r) {
        const nmsv_read& rd = reads[r];
        if (rd.seq >= n_seqs || lens[rd.seq] == 0) return set_err(ctx, NMSV_E_INVALID, "read %u: sequence index %u invalid or empty", r, rd.seq);
        if (rd.sv >= n_svs) return set_err(ctx, NMSV_E_INVALID, "read %u: SV index %u out of range", r, rd.sv);
        const nmsv_sv& sv = svs[rd.sv];
        if (r < sv.read_begin || r >= sv.read_end) return set_err(ctx, NMSV_E_INVALID, "read %u is outside the read range of its SV %u", r, rd.sv);
        const uint64_t n = lens[rd.seq];
        const bool sbnd = sv.flags & NMSV_SV_SBND;
        const bool mq = sbnd ? (rd.mapq1 != NMSV_MAPQ_NONE && rd.mapq1 >= sv.min_mapq)
                             : (rd.mapq1 != NMSV_MAPQ_NONE && rd.mapq1 >= sv.min_mapq && rd.mapq2 != NMSV_MAPQ_NONE && rd.mapq2 >= sv.min_mapq);
        const bool pruned = pl->prune && !mq;
        if (pruned) stt.reads_pruned += 1;
        else if (offsets[rd.seq] == NMSV_SEQ_ABSENT)
            return set_err(ctx, NMSV_E_INVALID, "read %u must be aligned but its sequence %u was not supplied (NMSV_SEQ_ABSENT)", r, rd.seq);
        uint64_t c = 0;
        bool spawner = false;
        for (int s = 0; s < 4; ++s) {
            if (sv.tmpl[s] == NMSV_NONE) continue;
            if (sbnd && (s == NMSV_SLOT_VAR2 || s == NMSV_SLOT_REF2)) continue;
            const uint64_t m = lens[sv.tmpl[s]];
            stt.cells_reference += 2 * m * n;
            int alias = s;
            for (int q = s - 1; q >= 0; --q) if (sv.tmpl[q] == sv.tmpl[s] && sv.tmpl_rc[q] == sv.tmpl_rc[s]) alias = q;
            if (alias != s) continue;
            const int cls = choose_class((uint32_t)m);
            const uint32_t nt = tiles_of((uint32_t)m, cls);
            stt.cells_dedup += 2 * m * n;
            if (nt > (uint32_t)kW) max_cols_multi = std::max<uint32_t>(max_cols_multi, (uint32_t)n);
            if (pruned) continue;
            if (pl->lazy_refs && s >= NMSV_SLOT_REF1) { spawner = true; continue; }      // queued on the device, if at all
            stt.cells_padded += 2ull * nt * 32 * kClassR[cls] * (n + 31);
            stt.fwd_tasks += 1;
            c += m * n;
        }
        // reads whose continuation can queue further forward work go first (their chain is three tasks deep), then
        // longest first: the tail of the launch is then made of short, independent tasks
        static const bool spawn_first = [] { const char* e = getenv("NMSV_SPAWN_FIRST"); return !(e && e[0] == '0'); }();   // measurement knob
        cost[r] = c | (spawner && c && spawn_first ? (1ull << 62) : 0ull);
    }
    std::vector<uint32_t> order(n_reads);
    for (uint32_t r = 0; r < n_reads; ++r) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });

    // ---- upload
    cudaStream_t s = ctx->stream;
    rc = upload_seq_classes(ctx, pl->st, n_seqs);
    if (rc) return rc;
    CUDA_TRY(ctx, pl->svs.alloc(sizeof(nmsv_sv) * n_svs, s));
    CUDA_TRY(ctx, pl->reads.alloc(sizeof(nmsv_read) * n_reads, s));
    CUDA_TRY(ctx, pl->order.alloc(sizeof(uint32_t) * n_reads, s));
    if (n_svs) CUDA_TRY(ctx, cudaMemcpyAsync(pl->svs.p, svs, sizeof(nmsv_sv) * n_svs, cudaMemcpyHostToDevice, s));
    if (n_reads) {
        CUDA_TRY(ctx, cudaMemcpyAsync(pl->reads.p, reads, sizeof(nmsv_read) * n_reads, cudaMemcpyHostToDevice, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(pl->order.p, order.data(), sizeof(uint32_t) * n_reads, cudaMemcpyHostToDevice, s));
    }
    stt.h2d_bytes = pl->st.h2d + sizeof(nmsv_sv) * (uint64_t)n_svs + (sizeof(nmsv_read) + 4) * (uint64_t)n_reads;

    // ---- workspace
    const uint64_t nslots = (uint64_t)n_reads * 4;
    CUDA_TRY(ctx, pl->tasks.alloc(sizeof(SwTask) * nslots, s));
    CUDA_TRY(ctx, pl->fwd.alloc(sizeof(SwResult) * nslots, s));
    const uint64_t n_regions = std::max<size_t>(1, pl->classes.size());
    if (n_regions > 8) return set_err(ctx, NMSV_E_INTERNAL, "more than 8 strip classes in one plan");
    // per read at most 4 begin tasks and 2 queued reference-allele tasks
    if ((uint64_t)n_reads * 6 > 0x7fffffffull) return set_err(ctx, NMSV_E_INVALID, "too many reads in one batch (%u)", n_reads);
    pl->rev_capacity = n_reads * 6u;
    CUDA_TRY(ctx, pl->rev_tasks.alloc(sizeof(SwTask) * (uint64_t)pl->rev_capacity * n_regions, s));
    CUDA_TRY(ctx, pl->rev_ready.alloc(sizeof(uint32_t) * (uint64_t)pl->rev_capacity * n_regions, s));
    CUDA_TRY(ctx, pl->group_remaining.alloc(sizeof(uint32_t) * (uint64_t)n_reads, s));
    CUDA_TRY(ctx, pl->hook.alloc(sizeof(GroupHook), s));
    CUDA_TRY(ctx, pl->cells.alloc(sizeof(unsigned long long) * kCellsTotal, s));
    CUDA_TRY(ctx, pl->rev.alloc(sizeof(SwResult) * nslots, s));
    CUDA_TRY(ctx, pl->sel.alloc(sizeof(Sel) * nslots, s));
    CUDA_TRY(ctx, pl->alns.alloc(sizeof(nmsv_aln) * nslots, s));
    CUDA_TRY(ctx, pl->flags.alloc(n_reads, s));
    CUDA_TRY(ctx, pl->supporting.alloc(sizeof(int32_t) * n_svs, s));
    CUDA_TRY(ctx, pl->records.alloc(sizeof(nmsv_record) * (uint64_t)n_reads, s));
    CUDA_TRY(ctx, pl->counters.alloc(sizeof(uint32_t) * kCntTotal, s));
    pl->bnd_stride = max_cols_multi ? max_cols_multi : 1;
    const uint64_t bnd_bytes = max_cols_multi ? (uint64_t)max_grid_gbufs(ctx) * pl->bnd_stride * sizeof(uint2) : 16;
    CUDA_TRY(ctx, pl->bnd.alloc(bnd_bytes, s));
    stt.workspace_bytes = pl->tasks.bytes + pl->fwd.bytes + pl->rev_tasks.bytes + pl->rev_ready.bytes +
                          pl->group_remaining.bytes + pl->rev.bytes + pl->sel.bytes +
                          pl->alns.bytes + pl->flags.bytes + pl->supporting.bytes + pl->records.bytes + pl->bnd.bytes;
    stt.n_classes = (uint32_t)pl->classes.size();
    stt.kernel_launches = (pl->fused_begin ? 2u + 5 * stt.n_classes : 3u + 2 * stt.n_classes);
    // the caller's host buffers may be reused as soon as we return
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    guard.p = nullptr;
    *out = pl;
    return NMSV_OK;
}

extern "C" int nmsv_plan_run(nmsv_plan* pl)
{
    if (!pl) return NMSV_E_INVALID;
    nmsv_ctx* ctx = pl->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t n_reads = pl->n_reads;
    CUDA_TRY(ctx, cudaMemsetAsync(pl->counters.p, 0, sizeof(uint32_t) * kCntTotal, s));
    CUDA_TRY(ctx, cudaMemsetAsync(pl->cells.p, 0, pl->cells.bytes, s));
    CUDA_TRY(ctx, cudaMemsetAsync(pl->supporting.p, 0, pl->supporting.bytes, s));
    pl->ran = true;
    if (n_reads == 0) return NMSV_OK;
    const int tpb = 128;
    const uint32_t nslots = n_reads * 4u;
    size_t ev = 0;
    pl->launch_phase.clear(); pl->launch_rclass.clear();
    auto note = [&](uint32_t phase, uint32_t rclass) { pl->launch_phase.push_back(phase); pl->launch_rclass.push_back(rclass); };
    mark(pl, ev, s);

    CUDA_TRY(ctx, cudaMemsetAsync(pl->group_remaining.p, 0, pl->group_remaining.bytes, s));
    CUDA_TRY(ctx, cudaMemsetAsync(pl->rev_ready.p, 0, pl->rev_ready.bytes, s));
    // reads that never reach a decision (pruned, or no task at all) must read as "no alignment, not a candidate"
    CUDA_TRY(ctx, cudaMemsetAsync(pl->flags.p, 0, pl->flags.bytes, s));
    CUDA_TRY(ctx, cudaMemsetAsync(pl->sel.p, 0, pl->sel.bytes, s));
    plan_fwd_kernel<<<(nslots + tpb - 1) / tpb, tpb, 0, s>>>(pl->svs.as<nmsv_sv>(), pl->reads.as<nmsv_read>(),
        pl->order.as<uint32_t>(), pl->st.offsets.as<uint64_t>(), pl->st.lens.as<uint32_t>(), pl->st.seq_class.as<uint8_t>(),
        n_reads, pl->tasks.as<SwTask>(), pl->group_remaining.as<uint32_t>(), pl->flags.as<uint8_t>(),
        pl->prune ? 1 : 0, pl->lazy_refs ? 1 : 0);
    note(kPhasePlan, 0); mark(pl, ev, s);

    GroupHook hk;
    memset(&hk, 0, sizeof(hk));
    DecideParams& dp = hk.d;
    dp.svs = pl->svs.as<nmsv_sv>(); dp.reads = pl->reads.as<nmsv_read>(); dp.offsets = pl->st.offsets.as<uint64_t>();
    dp.lens = pl->st.lens.as<uint32_t>(); dp.seq_class = pl->st.seq_class.as<uint8_t>(); dp.n_reads = n_reads;
    dp.fwd = pl->fwd.as<SwResult>(); dp.sel = pl->sel.as<Sel>(); dp.read_flags = pl->flags.as<uint8_t>();
    dp.rev_tasks = pl->rev_tasks.as<SwTask>(); dp.rev_ready = pl->rev_ready.as<uint32_t>(); dp.rev_capacity = pl->rev_capacity;
    dp.counters = pl->counters.as<uint32_t>();
    dp.group_remaining = pl->group_remaining.as<uint32_t>(); dp.cells = pl->cells.as<unsigned long long>();
    dp.lazy_refs = pl->lazy_refs ? 1 : 0;
    for (size_t i = 0; i < pl->classes.size(); ++i) dp.region_of_r[kClassR[pl->classes[i]]] = (uint8_t)i;
    dp.match = pl->sc.match; dp.open = pl->sc.gap_open; dp.ext = pl->sc.gap_extend;
    if (pl->fused_begin) {      // the hook lives in device memory; the host copy is pageable, so this copy is staged before returning
        CUDA_TRY(ctx, cudaMemcpyAsync(pl->hook.p, &hk, sizeof(hk), cudaMemcpyHostToDevice, s));
    }

    SwParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.codes = pl->st.codes.as<uint8_t>();
    sp.bnd = pl->bnd.as<uint2>(); sp.bnd_stride = pl->bnd_stride; sp.snap = (uint4*)ctx->snap;
    sp.rings = (uint2*)ctx->rings; sp.ring_cols = ctx->ring_cols; sp.ring_pos = (uint32_t*)ctx->ring_pos;
    for (sp.ring_shift = 0; (1u << sp.ring_shift) < ctx->ring_cols; ++sp.ring_shift) {}
    sp.match = pl->sc.match; sp.mismatch = pl->sc.mismatch; sp.open = pl->sc.gap_open; sp.ext = pl->sc.gap_extend;
    fill_packed(sp);
    sp.fwd_results = pl->fwd.as<SwResult>(); sp.rev_results = pl->rev.as<SwResult>();
    sp.cells = pl->cells.as<unsigned long long>();

    // Forward launches, one per strip class. Fused path: the warp that completes the last outstanding forward task of a
    // read runs its continuation (decide_read), which queues the read's reference-allele alignments or its begin tasks;
    // the launch of the tasks' class drains that queue alongside its static list.
    auto set_dyn = [&](size_t i) {
        sp.hook = pl->hook.as<GroupHook>(); sp.group_remaining = pl->group_remaining.as<uint32_t>();
        sp.dyn_tasks = dp.rev_tasks + i * (size_t)pl->rev_capacity; sp.dyn_ready = dp.rev_ready + i * (size_t)pl->rev_capacity;
        sp.dyn_head = pl->counters.as<uint32_t>() + kCntRevQueue + i;
        sp.dyn_reserve = pl->counters.as<uint32_t>() + kCntRevReserve + i;
    };
    sp.tasks = pl->tasks.as<SwTask>(); sp.n_tasks_dev = nullptr; sp.n_tasks = nslots;
    for (size_t i = 0; i < pl->classes.size(); ++i) {
        sp.queue_head = pl->counters.as<uint32_t>() + kCntFwdQueue + i;
        if (pl->fused_begin) set_dyn(i);
        launch_class_idx(pl->classes[i], sp, grid_for(ctx, pl->classes[i]), s);
        note(kPhaseFwd, (uint32_t)kClassR[pl->classes[i]]); mark(pl, ev, s);
    }

    if (pl->fused_begin) {
        // What is still queued: tasks whose class had its launch before they were queued. A chain is at most five
        // tasks deep (variant -> exact redo of the other variant segment -> first reference allele -> second -> begin
        // pass) and all speculative variant tasks are done by now, so four rounds see every task done. Each launch drains its
        // queue as a dynamic one (a continuation may append to the queue of the running launch); normally these launches
        // find nothing (a few microseconds each).
        sp.tasks = nullptr; sp.n_tasks_dev = nullptr; sp.n_tasks = 0;
        for (int round = 0; round < 4; ++round)
            for (size_t i = 0; i < pl->classes.size(); ++i) {
                sp.queue_head = pl->counters.as<uint32_t>() + kCntFwdQueue + i;
                set_dyn(i);
                launch_class_idx(pl->classes[i], sp, grid_for(ctx, pl->classes[i]), s);
                note(kPhaseRev, (uint32_t)kClassR[pl->classes[i]]); mark(pl, ev, s);
            }
    } else {
        decide_kernel<<<(n_reads + tpb - 1) / tpb, tpb, 0, s>>>(dp);
        note(kPhaseDecide, 0); mark(pl, ev, s);
        // begin launches over the queues as static lists
        sp.hook = nullptr; sp.group_remaining = nullptr;
        sp.dyn_tasks = nullptr; sp.dyn_ready = nullptr; sp.dyn_head = nullptr; sp.dyn_reserve = nullptr;
        sp.n_tasks = 0;
        for (size_t i = 0; i < pl->classes.size(); ++i) {
            sp.tasks = dp.rev_tasks + i * (size_t)pl->rev_capacity;
            sp.n_tasks_dev = pl->counters.as<uint32_t>() + kCntRevReserve + i;
            sp.queue_head = pl->counters.as<uint32_t>() + kCntRevQueue + i;
            launch_class_idx(pl->classes[i], sp, grid_for(ctx, pl->classes[i]), s);
            note(kPhaseRev, (uint32_t)kClassR[pl->classes[i]]); mark(pl, ev, s);
        }
    }

    FinalParams fp;
    fp.svs = dp.svs; fp.reads = dp.reads; fp.lens = dp.lens; fp.n_reads = n_reads; fp.sel = dp.sel;
    fp.rev = pl->rev.as<SwResult>(); fp.read_flags = dp.read_flags; fp.alns = pl->alns.as<nmsv_aln>();
    fp.supporting = pl->supporting.as<int32_t>(); fp.records = pl->records.as<nmsv_record>();
    fp.record_capacity = n_reads; fp.counters = dp.counters;
    final_kernel<<<(n_reads + tpb - 1) / tpb, tpb, 0, s>>>(fp);
    note(kPhaseFinal, 0); mark(pl, ev, s);
    if (pl->profiling) pl->prof_runs += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return NMSV_OK;
}

extern "C" int nmsv_plan_set_profiling(nmsv_plan* pl, int on)
{
    if (!pl) return NMSV_E_INVALID;
    pl->profiling = on != 0;
    pl->prof_runs = 0;
    return NMSV_OK;
}

extern "C" int nmsv_plan_kernel_times(nmsv_plan* pl, float* ms, uint32_t* phase, uint32_t* rclass, uint32_t capacity,
                                      uint32_t* n_launches)
{
    if (!pl || !n_launches) return NMSV_E_INVALID;
    nmsv_ctx* ctx = pl->ctx;
    const uint32_t n = (uint32_t)pl->launch_phase.size();
    *n_launches = n;
    if (!pl->profiling || !pl->ran || pl->prof_runs == 0 || n + 1 > nmsv_plan::kProfSlots)
        return set_err(ctx, NMSV_E_INVALID, "nmsv_plan_kernel_times needs a run with profiling enabled");
    if (capacity < n) return set_err(ctx, NMSV_E_CAPACITY, "kernel time buffers hold %u entries, %u launches", capacity, n);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t runs = std::min(pl->prof_runs, nmsv_plan::kProfRuns);
    std::vector<double> sum(n, 0.0);
    for (uint32_t k = 0; k < runs; ++k) {
        const size_t base = (size_t)((pl->prof_runs - 1 - k) % nmsv_plan::kProfRuns) * nmsv_plan::kProfSlots;
        CUDA_TRY(ctx, cudaEventSynchronize(pl->events[base + n]));
        for (uint32_t i = 0; i < n; ++i) {
            float t = 0.f;
            CUDA_TRY(ctx, cudaEventElapsedTime(&t, pl->events[base + i], pl->events[base + i + 1]));
            sum[i] += t;
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        if (ms) ms[i] = (float)(sum[i] / runs);
        if (phase) phase[i] = pl->launch_phase[i];
        if (rclass) rclass[i] = pl->launch_rclass[i];
    }
    return NMSV_OK;
}

extern "C" int nmsv_plan_copy_counts_device(nmsv_plan* pl, void* d_supporting_int32)
{
    if (!pl || !d_supporting_int32) return NMSV_E_INVALID;
    nmsv_ctx* ctx = pl->ctx;
    if (!pl->ran) return set_err(ctx, NMSV_E_INVALID, "nmsv_plan_copy_counts_device before nmsv_plan_run");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (pl->n_svs)
        CUDA_TRY(ctx, cudaMemcpyAsync(d_supporting_int32, pl->supporting.p, sizeof(int32_t) * pl->n_svs,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return NMSV_OK;
}

extern "C" int nmsv_plan_download(nmsv_plan* pl, int32_t* checked, int32_t* supporting, nmsv_record* records,
                                  uint64_t record_capacity, uint64_t* n_records)
{
    if (!pl) return NMSV_E_INVALID;
    nmsv_ctx* ctx = pl->ctx;
    if (!pl->ran) return set_err(ctx, NMSV_E_INVALID, "nmsv_plan_download before nmsv_plan_run");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint32_t cnt[kCntTotal];
    unsigned long long cells[kCellsTotal];
    CUDA_TRY(ctx, cudaMemcpyAsync(cnt, pl->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(cells, pl->cells.p, sizeof(cells), cudaMemcpyDeviceToHost, s));
    if (supporting && pl->n_svs)
        CUDA_TRY(ctx, cudaMemcpyAsync(supporting, pl->supporting.p, sizeof(int32_t) * pl->n_svs, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    uint64_t queued = 0;
    for (uint32_t i = 0; i < 8; ++i) queued += cnt[kCntRevReserve + i];
    pl->stats.rev_tasks = queued - cells[kTasksSpawnedFwd];
    pl->stats.fwd_tasks_queued = cells[kTasksSpawnedFwd];
    pl->stats.cells_executed = pl->n_reads ? cells[kCellsFwd] : 0;
    pl->stats.cells_begin = cells[kCellsBegin];
    pl->stats.cells_pruned = pl->stats.cells_dedup - pl->stats.cells_executed;
    pl->stats.d2h_bytes = sizeof(cnt) + sizeof(cells) + sizeof(int32_t) * (uint64_t)pl->n_svs;
    if (cnt[kCntSaturated])
        return set_err(ctx, NMSV_E_RANGE, "%u reads have an alignment whose score leaves the 16-bit range: parasail.ssw returns "
                       "None there and the reference's ssw_check fails (count_sread_by_alignment.py:151)", cnt[kCntSaturated]);
    if (cnt[kCntStatus])
        return set_err(ctx, NMSV_E_INTERNAL, "%u begin-pass windows did not reproduce their forward score", cnt[kCntStatus]);
    if (checked) for (uint32_t i = 0; i < pl->n_svs; ++i) checked[i] = pl->checked[i];
    const uint64_t nrec = cnt[kCntRecords];
    if (n_records) *n_records = nrec;
    if (records) {
        if (nrec > record_capacity)
            return set_err(ctx, NMSV_E_CAPACITY, "record buffer holds %llu entries but %llu reads are supporting",
                           (unsigned long long)record_capacity, (unsigned long long)nrec);
        if (nrec) {
            CUDA_TRY(ctx, cudaMemcpyAsync(records, pl->records.p, sizeof(nmsv_record) * nrec, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(ctx, cudaStreamSynchronize(s));
            pl->stats.d2h_bytes += sizeof(nmsv_record) * nrec;
        }
    }
    return NMSV_OK;
}

extern "C" int nmsv_plan_download_alns(nmsv_plan* pl, nmsv_aln* alns, uint8_t* read_flags)
{
    if (!pl) return NMSV_E_INVALID;
    nmsv_ctx* ctx = pl->ctx;
    if (!pl->ran) return set_err(ctx, NMSV_E_INVALID, "nmsv_plan_download_alns before nmsv_plan_run");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (alns && pl->n_reads)
        CUDA_TRY(ctx, cudaMemcpyAsync(alns, pl->alns.p, sizeof(nmsv_aln) * 4ull * pl->n_reads, cudaMemcpyDeviceToHost, s));
    if (read_flags && pl->n_reads)
        CUDA_TRY(ctx, cudaMemcpyAsync(read_flags, pl->flags.p, pl->n_reads, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return NMSV_OK;
}

extern "C" int nmsv_plan_stats(const nmsv_plan* pl, nmsv_stats* stats)
{
    if (!pl || !stats) return NMSV_E_INVALID;
    *stats = pl->stats;
    return NMSV_OK;
}

extern "C" int nmsv_validate_batch(nmsv_ctx* ctx, const uint8_t* codes, uint64_t n_codes, const uint64_t* offsets,
                                   const uint32_t* lens, uint32_t n_seqs, const nmsv_sv* svs, uint32_t n_svs,
                                   const nmsv_read* reads, uint32_t n_reads, const nmsv_scoring* scoring,
                                   int32_t* checked, int32_t* supporting, nmsv_record* records,
                                   uint64_t record_capacity, uint64_t* n_records)
{
    static const bool trace = getenv("NMSV_TRACE") != nullptr;      // host-side phase times on stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    nmsv_plan* pl = nullptr;
    int rc = nmsv_plan_create(ctx, &pl, codes, n_codes, offsets, lens, n_seqs, svs, n_svs, reads, n_reads, scoring);
    if (rc) return rc;
    const double t1 = trace ? now() : 0.0;
    rc = nmsv_plan_run(pl);
    const double t2 = trace ? now() : 0.0;
    if (!rc) rc = nmsv_plan_download(pl, checked, supporting, records, record_capacity, n_records);
    const double t3 = trace ? now() : 0.0;
    nmsv_plan_destroy(pl);
    if (trace)
        fprintf(stderr, "[nmsv] validate_batch: create+upload %.2f ms, launch %.2f ms, wait+download %.2f ms, destroy %.2f ms\n",
                t1 - t0, t2 - t1, t3 - t2, now() - t3);
    return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// pair APIs (operator seam and function seam)
// ------------------------------------------------------------------------------------------------------------------
static int run_pairs(nmsv_ctx* ctx, const uint8_t* codes, uint64_t n_codes, const uint64_t* offsets,
                     const uint32_t* lens, uint32_t n_seqs, const uint32_t* tmpl_idx, const uint32_t* rc_idx,
                     const uint32_t* read_idx, uint64_t n_pairs64, const nmsv_scoring* sc, int both_strands,
                     int want_begin, nmsv_ssw_result* out_ssw, nmsv_aln* out_aln)
{
    if (!ctx) return NMSV_E_INVALID;
    int rc = check_scoring(ctx, sc);
    if (rc) return rc;
    rc = check_seq_tables(ctx, codes, n_codes, offsets, lens, n_seqs);
    if (rc) return rc;
    if (n_pairs64 == 0) return NMSV_OK;
    if (!tmpl_idx || !read_idx || (!out_ssw && !out_aln)) return set_err(ctx, NMSV_E_INVALID, "pair index / output arrays must not be NULL");
    if (n_pairs64 > 0x7fffffffu) return set_err(ctx, NMSV_E_INVALID, "too many pairs in one call");
    const uint32_t n_pairs = (uint32_t)n_pairs64;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));

    SeqTables st;
    // the sequence upload runs while the host checks the pairs; an error return waits for it (the caller's buffers)
    struct InFlight { cudaStream_t s; bool armed; ~InFlight() { if (armed) cudaStreamSynchronize(s); } } in_flight{ctx->stream, true};
    rc = upload_seq_codes(ctx, st, codes, n_codes, offsets, lens, n_seqs);
    if (rc) return rc;
    st.cls_host.assign(n_seqs, 0);
    bool class_used[8] = {false};
    uint32_t max_cols_multi = 0;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        rc = check_template(ctx, tmpl_idx[i], offsets, lens, n_seqs, sc, "template");
        if (rc) return rc;
        if (rc_idx && rc_idx[i] != NMSV_NONE) {
            rc = check_template(ctx, rc_idx[i], offsets, lens, n_seqs, sc, "reverse-complement template");
            if (rc) return rc;
            if (lens[rc_idx[i]] != lens[tmpl_idx[i]]) return set_err(ctx, NMSV_E_INVALID, "pair %u: explicit reverse complement has a different length", i);
        }
        if (read_idx[i] >= n_seqs || lens[read_idx[i]] == 0 || offsets[read_idx[i]] == NMSV_SEQ_ABSENT)
            return set_err(ctx, NMSV_E_INVALID, "pair %u: read index %u invalid or empty", i, read_idx[i]);
        const int c = choose_class(lens[tmpl_idx[i]]);
        st.cls_host[tmpl_idx[i]] = (uint8_t)kClassR[c];
        class_used[c] = true;
        if (tiles_of(lens[tmpl_idx[i]], c) > (uint32_t)kW) max_cols_multi = std::max(max_cols_multi, lens[read_idx[i]]);
    }
    std::vector<int> classes;
    for (int c = 0; c < kNumClasses; ++c) if (class_used[c]) classes.push_back(c);

    cudaStream_t s = ctx->stream;
    rc = upload_seq_classes(ctx, st, n_seqs);
    if (rc) return rc;
    DevBuf d_t, d_rc, d_r, tasks, fwd, rev_tasks, rev_ready, rev, sel, counters, bnd, d_out;
    CUDA_TRY(ctx, d_t.alloc(4ull * n_pairs, s));
    CUDA_TRY(ctx, d_r.alloc(4ull * n_pairs, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_t.p, tmpl_idx, 4ull * n_pairs, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_r.p, read_idx, 4ull * n_pairs, cudaMemcpyHostToDevice, s));
    if (rc_idx) {
        CUDA_TRY(ctx, d_rc.alloc(4ull * n_pairs, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_rc.p, rc_idx, 4ull * n_pairs, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(ctx, tasks.alloc(sizeof(SwTask) * (uint64_t)n_pairs, s));
    CUDA_TRY(ctx, fwd.alloc(sizeof(SwResult) * (uint64_t)n_pairs, s));
    const uint64_t n_regions = want_begin ? std::max<size_t>(1, classes.size()) : 1;
    const uint64_t rev_cap = want_begin ? n_pairs : 1;
    CUDA_TRY(ctx, rev_tasks.alloc(sizeof(SwTask) * rev_cap * n_regions, s));
    CUDA_TRY(ctx, rev_ready.alloc(sizeof(uint32_t) * rev_cap * n_regions, s));
    CUDA_TRY(ctx, rev.alloc(sizeof(SwResult) * (uint64_t)n_pairs, s));
    CUDA_TRY(ctx, sel.alloc(sizeof(Sel) * (uint64_t)n_pairs, s));
    CUDA_TRY(ctx, counters.alloc(sizeof(uint32_t) * kCntTotal, s));
    CUDA_TRY(ctx, cudaMemsetAsync(counters.p, 0, sizeof(uint32_t) * kCntTotal, s));
    const uint32_t stride = max_cols_multi ? max_cols_multi : 1;
    CUDA_TRY(ctx, bnd.alloc(max_cols_multi ? (uint64_t)max_grid_gbufs(ctx) * stride * sizeof(uint2) : 16, s));
    const size_t out_bytes = out_ssw ? sizeof(nmsv_ssw_result) * (size_t)n_pairs : sizeof(nmsv_aln) * (size_t)n_pairs;
    CUDA_TRY(ctx, d_out.alloc(out_bytes, s));

    PairParams pp;
    pp.offsets = st.offsets.as<uint64_t>(); pp.lens = st.lens.as<uint32_t>(); pp.seq_class = st.seq_class.as<uint8_t>();
    pp.tmpl_idx = d_t.as<uint32_t>(); pp.rc_idx = rc_idx ? d_rc.as<uint32_t>() : nullptr; pp.read_idx = d_r.as<uint32_t>();
    pp.n_pairs = n_pairs; pp.both_strands = both_strands;
    pp.fwd = fwd.as<SwResult>(); pp.sel = sel.as<Sel>(); pp.rev_tasks = rev_tasks.as<SwTask>(); pp.rev = rev.as<SwResult>();
    pp.rev_ready = rev_ready.as<uint32_t>(); pp.rev_capacity = (uint32_t)rev_cap;
    memset(pp.region_of_r, 0, sizeof(pp.region_of_r));
    for (size_t i = 0; i < classes.size(); ++i) pp.region_of_r[kClassR[classes[i]]] = (uint8_t)i;
    pp.counters = counters.as<uint32_t>();
    pp.out_ssw = out_ssw ? d_out.as<nmsv_ssw_result>() : nullptr;
    pp.out_aln = out_ssw ? nullptr : d_out.as<nmsv_aln>();
    pp.want_begin = want_begin; pp.match = sc->match; pp.open = sc->gap_open; pp.ext = sc->gap_extend;

    const int tpb = 128;
    const int nb = (int)((n_pairs + tpb - 1) / tpb);
    pair_plan_kernel<<<nb, tpb, 0, s>>>(pp, tasks.as<SwTask>());
    SwParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.codes = st.codes.as<uint8_t>(); sp.bnd = bnd.as<uint2>(); sp.bnd_stride = stride; sp.snap = (uint4*)ctx->snap;
    sp.rings = (uint2*)ctx->rings; sp.ring_cols = ctx->ring_cols; sp.ring_pos = (uint32_t*)ctx->ring_pos;
    for (sp.ring_shift = 0; (1u << sp.ring_shift) < ctx->ring_cols; ++sp.ring_shift) {}
    sp.match = sc->match; sp.mismatch = sc->mismatch; sp.open = sc->gap_open; sp.ext = sc->gap_extend;
    fill_packed(sp);
    sp.tasks = tasks.as<SwTask>(); sp.n_tasks_dev = nullptr; sp.n_tasks = n_pairs;
    sp.fwd_results = fwd.as<SwResult>(); sp.rev_results = rev.as<SwResult>();
    for (size_t i = 0; i < classes.size(); ++i) {
        sp.queue_head = counters.as<uint32_t>() + kCntFwdQueue + i;
        launch_class_idx(classes[i], sp, grid_for(ctx, classes[i]), s);
    }
    pair_decide_kernel<<<nb, tpb, 0, s>>>(pp);
    if (want_begin) {
        sp.n_tasks = 0;
        for (size_t i = 0; i < classes.size(); ++i) {
            sp.tasks = rev_tasks.as<SwTask>() + i * rev_cap;
            sp.n_tasks_dev = counters.as<uint32_t>() + kCntRevReserve + i;
            sp.queue_head = counters.as<uint32_t>() + kCntRevQueue + i;
            launch_class_idx(classes[i], sp, grid_for(ctx, classes[i]), s);
        }
    }
    pair_final_kernel<<<nb, tpb, 0, s>>>(pp);
    CUDA_TRY(ctx, cudaGetLastError());
    uint32_t cnt[kCntTotal];
    CUDA_TRY(ctx, cudaMemcpyAsync(cnt, counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_ssw ? (void*)out_ssw : (void*)out_aln, d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    if (cnt[kCntStatus])
        return set_err(ctx, NMSV_E_INTERNAL, "%u begin-pass windows did not reproduce their forward score", cnt[kCntStatus]);
    return NMSV_OK;
}

extern "C" int nmsv_ssw_batch(nmsv_ctx* ctx, const uint8_t* codes, uint64_t n_codes, const uint64_t* offsets,
                              const uint32_t* lens, uint32_t n_seqs, const uint32_t* tmpl_idx, const uint32_t* read_idx,
                              uint64_t n_pairs, const nmsv_scoring* scoring, int want_begin, nmsv_ssw_result* out)
{
    return run_pairs(ctx, codes, n_codes, offsets, lens, n_seqs, tmpl_idx, nullptr, read_idx, n_pairs, scoring, 0,
                     want_begin, out, nullptr);
}

extern "C" int nmsv_ssw_check_batch(nmsv_ctx* ctx, const uint8_t* codes, uint64_t n_codes, const uint64_t* offsets,
                                    const uint32_t* lens, uint32_t n_seqs, const uint32_t* tmpl_idx,
                                    const uint32_t* tmpl_rc_idx, const uint32_t* read_idx, uint64_t n_pairs,
                                    const nmsv_scoring* scoring, nmsv_aln* out)
{
    return run_pairs(ctx, codes, n_codes, offsets, lens, n_seqs, tmpl_idx, tmpl_rc_idx, read_idx, n_pairs, scoring, 1,
                     1, nullptr, out);
}


// ------------------------------------------------------------------------------------------------------------------
// "next" row: sw_jump (reference smith_waterman.py:5-127)
// ------------------------------------------------------------------------------------------------------------------
// One traceback written out backwards (reference smith_waterman.py:71-125): operation 1 consumes a base of both sequences,
// 2 a contig base against a gap, 3 a region base against a gap, 4 (the jump cell, H2 part only) shows both bases without moving.
static bool jump_walk(const uint8_t* c, uint32_t nc, const uint8_t* r, uint32_t nr, int32_t i, int32_t j, const uint8_t* ops,
                      uint32_t n, uint8_t* oc, uint8_t* og)
{
    for (uint32_t k = 0; k < n; ++k) {
        const uint8_t o = ops[k];
        const bool uc = o == 1 || o == 2 || o == 4, ur = o == 1 || o == 3 || o == 4;
        if ((uc && (i < 1 || (uint32_t)i > nc)) || (ur && (j < 1 || (uint32_t)j > nr)) || o < 1 || o > 4) return false;
        oc[n - 1 - k] = uc ? c[i - 1] : (uint8_t)'-';
        og[n - 1 - k] = ur ? r[j - 1] : (uint8_t)'-';
        if (o == 1 || o == 2) --i;
        if (o == 1 || o == 3) --j;
    }
    return true;
}

extern "C" int nmsv_sw_jump_strings(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* lens,
                                    const uint32_t* contig_idx, const uint32_t* region1_idx, const uint32_t* region2_idx,
                                    uint32_t n_problems, const nmsv_jump_result* results, const uint8_t* ops,
                                    const uint64_t* row_offsets, uint8_t* contig_rows, uint8_t* region_rows)
{
    if (n_problems == 0) return NMSV_OK;
    if (!bytes || !offsets || !lens || !contig_idx || !region1_idx || !region2_idx || !results || !ops || !row_offsets ||
        !contig_rows || !region_rows)
        return NMSV_E_INVALID;
    for (uint32_t p = 0; p < n_problems; ++p) {
        const nmsv_jump_result& r = results[p];
        if (!r.found) continue;
        const uint32_t ci = contig_idx[p], ai = region1_idx[p], bi = region2_idx[p];
        const uint64_t cap = 2ull * lens[ci] + lens[ai] + lens[bi] + 4ull;
        if ((uint64_t)r.n_ops1 + r.n_ops2 + 1ull > cap) return NMSV_E_INVALID;
        const uint8_t* o = ops + row_offsets[p];
        uint8_t* oc = contig_rows + row_offsets[p];
        uint8_t* og = region_rows + row_offsets[p];
        // "<H1 part> <H2 part>": the operations hold the H2 part first (the traceback starts there)
        if (!jump_walk(bytes + offsets[ci], lens[ci], bytes + offsets[ai], lens[ai], r.i1_end, r.j1_end, o + r.n_ops2, r.n_ops1, oc, og))
            return NMSV_E_INVALID;
        oc[r.n_ops1] = (uint8_t)' ';
        og[r.n_ops1] = (uint8_t)' ';
        if (!jump_walk(bytes + offsets[ci], lens[ci], bytes + offsets[bi], lens[bi], r.i2_end, r.j2_end, o, r.n_ops2,
                       oc + r.n_ops1 + 1, og + r.n_ops1 + 1))
            return NMSV_E_INVALID;
    }
    return NMSV_OK;
}

extern "C" int nmsv_sw_jump_batch(nmsv_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* offsets,
                                  const uint32_t* lens, uint32_t n_seqs, const uint32_t* contig_idx,
                                  const uint32_t* region1_idx, const uint32_t* region2_idx, uint32_t n_problems,
                                  const nmsv_jump_params* params, const double* log_table, uint32_t log_table_len,
                                  nmsv_jump_result* out, uint8_t* ops, uint64_t ops_capacity)
{
    static_assert(sizeof(nmsv_jump_result) == sizeof(JumpResult), "result layouts must match");
    if (!ctx) return NMSV_E_INVALID;
    if (n_problems == 0) return NMSV_OK;
    if (!params || !contig_idx || !region1_idx || !region2_idx || !out || !ops)
        return set_err(ctx, NMSV_E_INVALID, "nmsv_sw_jump_batch: NULL argument");
    int rc = check_seq_tables(ctx, bytes, n_bytes, offsets, lens, n_seqs);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    static const bool jtrace_all = getenv("NMSV_TRACE") != nullptr;
    const auto ja0 = std::chrono::steady_clock::now();

    // Regions of up to kJumpWarpMaxRegion columns (every locate_bp call: +-20 bp around a read-supported breakpoint range, or
    // the first / last 500 bases of a contig) run one warp per problem; longer ones keep one block per problem.
    std::vector<JumpProblem> probs;
    std::vector<JumpWarpProblem> wprobs;
    uint64_t ops_total = 0;
    uint32_t max_region = 0;
    for (uint32_t i = 0; i < n_problems; ++i) {
        const uint32_t c = contig_idx[i], a = region1_idx[i], b = region2_idx[i];
        if (c >= n_seqs || a >= n_seqs || b >= n_seqs)
            return set_err(ctx, NMSV_E_INVALID, "sw_jump problem %u: sequence index out of range", i);
        if (lens[a] > (uint32_t)kJumpMaxRegion || lens[b] > (uint32_t)kJumpMaxRegion)
            return set_err(ctx, NMSV_E_RANGE, "sw_jump problem %u: region longer than %d", i, kJumpMaxRegion);
        // the warp kernel keeps its values times 32 in 32-bit lanes: it takes the problems whose scores stay below 2^23
        // (every real one: a few thousand bases times a score of 1..3); anything else runs on the block kernel
        const uint64_t unit = (uint64_t)std::max(std::max(std::abs((long long)params->match_score), std::abs((long long)params->mismatch_penalty)),
                                                 std::abs((long long)params->gap_cost)) + 1;
        const uint64_t bound = ((uint64_t)lens[c] + lens[a] + lens[b] + 4) * unit + 64ull * (uint64_t)std::abs((long long)params->jump_cost) +
                               (uint64_t)std::abs((long long)params->minimum_score);
        if (std::max(lens[a], lens[b]) <= (uint32_t)kJumpWarpMaxRegion && bound < (1ull << 23)) {
            JumpWarpProblem p;
            p.c_off = offsets[c]; p.r1_off = offsets[a]; p.r2_off = offsets[b];
            p.n = lens[c]; p.m1 = lens[a]; p.m2 = lens[b]; p.res_idx = i;
            p.w1_off = p.w2_off = p.row_off = 0;
            p.ops_off = ops_total;
            wprobs.push_back(p);
        } else {
            JumpProblem p;
            p.c_off = offsets[c]; p.r1_off = offsets[a]; p.r2_off = offsets[b];
            p.n = lens[c]; p.m1 = lens[a]; p.m2 = lens[b]; p.res_idx = i;
            p.ops_off = ops_total;
            max_region = std::max(max_region, std::max(p.m1, p.m2));
            probs.push_back(p);
        }
        ops_total += 2ull * lens[c] + lens[a] + lens[b] + 4;
    }
    if (ops_total > ops_capacity)
        return set_err(ctx, NMSV_E_CAPACITY, "ops buffer holds %llu bytes, %llu needed", (unsigned long long)ops_capacity,
                       (unsigned long long)ops_total);

    DevBuf d_bytes, d_log, d_probs, d_paths, d_rows, d_res, d_ops, d_next;
    CUDA_TRY(ctx, d_bytes.alloc(n_bytes, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_bytes.p, bytes, n_bytes, cudaMemcpyHostToDevice, s));
    if (log_table && log_table_len) {
        CUDA_TRY(ctx, d_log.alloc(sizeof(double) * log_table_len, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_log.p, log_table, sizeof(double) * log_table_len, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(ctx, d_res.alloc(sizeof(JumpResult) * (size_t)n_problems, s));
    CUDA_TRY(ctx, d_ops.alloc(ops_total, s));
    const uint64_t kMaxWorkBytes = 6ull << 30;       // workspace per launch

    // ---- warp-per-problem kernel, longest problems first, in chunks whose workspace stays below the limit
    std::stable_sort(wprobs.begin(), wprobs.end(), [](const JumpWarpProblem& x, const JumpWarpProblem& y) {
        return (uint64_t)x.n * (x.m1 + x.m2) > (uint64_t)y.n * (y.m1 + y.m2); });
    for (size_t first = 0; first < wprobs.size();) {
        uint64_t words = 0;
        size_t last = first;
        while (last < wprobs.size()) {
            JumpWarpProblem& p = wprobs[last];
            const uint64_t rows = (uint64_t)p.n + 1, wsteps = (uint64_t)p.n + 32;      // plane words are indexed by (step, lane)
            const uint64_t need = 64 * wsteps + 2 * rows + 2 + 2 * rows;      // two word matrices, rowbest, prefix (+pad), rowinfo
            if (last > first && (words + need) * 8 > kMaxWorkBytes) break;
            p.w1_off = words; p.w2_off = words + 32 * wsteps; p.row_off = words + 64 * wsteps;
            words += need;
            ++last;
        }
        const uint32_t cnt = (uint32_t)(last - first);
        CUDA_TRY(ctx, d_rows.alloc(words * 8, s));
        CUDA_TRY(ctx, d_probs.alloc(sizeof(JumpWarpProblem) * (size_t)cnt, s));
        CUDA_TRY(ctx, d_next.alloc(sizeof(uint32_t), s));
        CUDA_TRY(ctx, cudaMemsetAsync(d_next.p, 0, sizeof(uint32_t), s));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_probs.p, wprobs.data() + first, sizeof(JumpWarpProblem) * (size_t)cnt, cudaMemcpyHostToDevice, s));
        JumpWarpParams wp;
        wp.bytes = d_bytes.as<uint8_t>(); wp.problems = d_probs.as<JumpWarpProblem>(); wp.n_problems = cnt;
        wp.next = d_next.as<uint32_t>(); wp.work = d_rows.as<unsigned long long>();
        wp.logtab = d_log.as<double>(); wp.logtab_len = (log_table && log_table_len) ? log_table_len : 0;
        wp.results = d_res.as<JumpResult>(); wp.ops = d_ops.as<uint8_t>();
        wp.match = params->match_score; wp.mismatch_penalty = params->mismatch_penalty; wp.gap_cost = params->gap_cost;
        wp.jump_cost = params->jump_cost; wp.minimum_score = params->minimum_score;
        // Two blocks = eight warps per SM, on purpose: the longest problems of a batch (a 2.6 kb contig against two 900-base
        // regions is 5 200 steps of ~30 cells) set the length of the launch, and a warp that shares its scheduler with seven
        // others walks them at an eighth of the issue rate. Measured on the 3000-problem set: 1 / 2 / 3 / 4 / 8 blocks per SM
        // 9.6 / 6.2 / 6.3 / 8.1 / 9.4 ms (two warps per scheduler already issue ~0.7 instructions per clock).
        static const uint32_t jbps = getenv("NMSV_JUMP_BPS") ? (uint32_t)std::max(1, atoi(getenv("NMSV_JUMP_BPS"))) : 2u;   // tuning knob
        const uint32_t blocks = std::min<uint32_t>((cnt + kJumpWarpsPerBlock - 1) / kJumpWarpsPerBlock,
                                                   (uint32_t)ctx->prop.multiProcessorCount * jbps);
        static const bool jtrace = getenv("NMSV_TRACE") != nullptr;
        const auto jt0 = std::chrono::steady_clock::now();
        if (jtrace) cudaStreamSynchronize(s);
        const auto jt1 = std::chrono::steady_clock::now();
        sw_jump_warp_kernel<<<blocks, kJumpWarpsPerBlock * 32, kJumpWarpsPerBlock * kJumpWarpSmem, s>>>(wp);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaStreamSynchronize(s));      // the host-side problem table of this chunk may be reused
        if (jtrace)
            fprintf(stderr, "[nmsv] sw_jump: %u problems, workspace %.1f MB, setup %.2f ms, warp kernel %.2f ms\n", cnt, words * 8 / 1e6,
                    std::chrono::duration<double, std::milli>(jt1 - jt0).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - jt1).count());
        first = last;
    }

    // ---- block-per-problem kernel for the long regions
    const size_t smem = 3ull * (max_region + 2) * sizeof(int32_t);
    int bps = 1;
    if (!probs.empty())
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sw_jump_kernel, kJumpThreads, smem));
    const uint32_t max_blocks = (uint32_t)ctx->prop.multiProcessorCount * (uint32_t)std::max(1, bps);
    const uint32_t n_slow = (uint32_t)probs.size();
    uint32_t first = 0;
    while (first < n_slow) {
        uint64_t path_bytes = 0, row_words = 0;
        uint32_t last = first;
        while (last < n_slow) {
            JumpProblem& p = probs[last];
            const uint64_t pb = (uint64_t)(p.n + 1) * (p.m1 + 1) + (uint64_t)(p.n + 1) * (p.m2 + 1);
            if (last > first && path_bytes + pb > kMaxWorkBytes) break;
            p.p1_off = path_bytes;
            p.p2_off = path_bytes + (uint64_t)(p.n + 1) * (p.m1 + 1);
            path_bytes += pb;
            p.row_off = row_words;
            row_words += 3ull * (p.n + 1) + ((uint64_t)p.n + p.m2 + 3) / 2 + 1;
            ++last;
        }
        const uint32_t cnt = last - first;
        CUDA_TRY(ctx, d_paths.alloc(path_bytes, s));
        CUDA_TRY(ctx, d_rows.alloc(row_words * 8, s));
        CUDA_TRY(ctx, d_probs.alloc(sizeof(JumpProblem) * (size_t)cnt, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_probs.p, probs.data() + first, sizeof(JumpProblem) * (size_t)cnt, cudaMemcpyHostToDevice, s));
        JumpParams jp;
        jp.bytes = d_bytes.as<uint8_t>(); jp.problems = d_probs.as<JumpProblem>(); jp.n_problems = cnt;
        jp.paths = d_paths.as<uint8_t>(); jp.rows = d_rows.as<unsigned long long>();
        jp.logtab = d_log.as<double>(); jp.logtab_len = (log_table && log_table_len) ? log_table_len : 0;
        jp.results = d_res.as<JumpResult>(); jp.ops = d_ops.as<uint8_t>();
        jp.match = params->match_score; jp.mismatch_penalty = params->mismatch_penalty; jp.gap_cost = params->gap_cost;
        jp.jump_cost = params->jump_cost; jp.minimum_score = params->minimum_score;
        sw_jump_kernel<<<std::min(cnt, max_blocks), kJumpThreads, smem, s>>>(jp);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaStreamSynchronize(s));      // the host-side problem table of this chunk may be reused
        first = last;
    }
    const auto ja1 = std::chrono::steady_clock::now();
    // The operations come back through a pinned staging buffer (a pageable destination moved 11 MB at 3.3 GB/s: 3.4 ms of a
    // 12 ms call), and only the operations actually written -- about a quarter of the upper-bound layout, the path ends at
    // the jump -- are copied on into the caller's array.
    if (ops_total > ctx->jump_stage_bytes) {
        if (ctx->jump_stage) cudaFreeHost(ctx->jump_stage);
        ctx->jump_stage = nullptr; ctx->jump_stage_bytes = 0;
        const size_t cap = std::max<size_t>(ops_total + ops_total / 4, (size_t)1 << 20);
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->jump_stage, cap, cudaHostAllocDefault));
        ctx->jump_stage_bytes = cap;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out, d_res.p, sizeof(JumpResult) * (size_t)n_problems, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->jump_stage, d_ops.p, ops_total, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    {
        uint64_t off = 0;
        const uint8_t* stage = static_cast<const uint8_t*>(ctx->jump_stage);
        for (uint32_t i = 0; i < n_problems; ++i) {
            const size_t cnt = (size_t)out[i].n_ops2 + out[i].n_ops1;
            if (cnt) memcpy(ops + off, stage + off, cnt);
            off += 2ull * lens[contig_idx[i]] + lens[region1_idx[i]] + lens[region2_idx[i]] + 4;
        }
    }
    if (jtrace_all)
        fprintf(stderr, "[nmsv] sw_jump: whole call up to the kernels' end %.2f ms, copy back of %.1f MB of results and operations %.2f ms\n",
                std::chrono::duration<double, std::milli>(ja1 - ja0).count(), (ops_total + sizeof(JumpResult) * (double)n_problems) / 1e6,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ja1).count());
    return NMSV_OK;
}
