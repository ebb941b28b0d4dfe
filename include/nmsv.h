/*
 * nmsv.h -- C ABI of libnmsv_sw.so: B200 (sm_100a) engine for nanomonsv's supporting-read validation path.
 *
 * Plain C linkage, plain pointers and sizes, caller-owned buffers; loadable with ctypes / cgo / JNI alike.
 * Every entry point replaces one seam of the reference (friend1ws/nanomonsv v0.8.0, paths relative to the
 * reference checkout):
 *
 *   nmsv_ssw_batch        <- parasail.ssw(s1, s2, open, extend, matrix)            count_sread_by_alignment.py:148-149
 *                            (also post_proc.py:118-120, locate_bp_sbnd.py:28-29, generate_consensus.py:119)
 *   nmsv_ssw_check_batch  <- ssw_check(query_fa, target_fa, use_ssw_lib)           count_sread_by_alignment.py:137-164
 *                            (legacy twin long_read_validate.py:12-17,126-153)
 *   nmsv_validate_batch   <- Alignment_counter.count_alignment_and_print[_sbnd]   count_sread_by_alignment.py:292-397
 *                            i.e. the arithmetic of count_sread_by_alignment()     count_sread_by_alignment.py:401-451
 *   nmsv_plan_* / nmsv_plan_run / nmsv_plan_download: the same stage split into upload / device-resident run /
 *                            download so a caller (bench.py) can time the device step on its own.
 *   The ctypes precedent in the reference is ssw_lib.CSsw (ssw_lib.py:95-199): unlike it, nothing returned by
 *   this library has to be freed by the caller except the handles created by nmsv_create / nmsv_plan_create.
 *
 * Sequence encoding (all APIs): one byte per base; 0,1,2,3 = A,C,G,T (case folded by the caller), 4 = wildcard
 * (any other byte; scores 0 against everything, like the extra row/column of parasail.matrix_create), see
 * nmsv_encode_ascii.  All sequences of a call live in ONE buffer `codes`; sequence s occupies
 * codes[offsets[s] .. offsets[s] + lens[s]).
 *
 * Error convention: every int function returns NMSV_OK (0) or a negative NMSV_E_* code; the message is available
 * from nmsv_last_error(ctx).  There is no CPU fallback: without a usable CUDA device nmsv_create fails.
 */
#ifndef NMSV_H
#define NMSV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMSV_ABI_VERSION 2

#define NMSV_OK 0
#define NMSV_E_INVALID (-1)   /* bad argument (NULL pointer, empty sequence, index out of range ...) */
#define NMSV_E_CUDA (-2)      /* a CUDA runtime call failed */
#define NMSV_E_RANGE (-3)     /* scoring parameters too large for 16-bit lanes, or (stage seam) an alignment whose score
                                 left the 16-bit range: parasail.ssw returns None and the reference's ssw_check fails */
#define NMSV_E_CAPACITY (-4)  /* caller's record buffer too small; *n_records holds the required count */
#define NMSV_E_INTERNAL (-5)  /* device-side consistency check failed */

#define NMSV_NONE 0xFFFFFFFFu /* "no template in this slot" */
#define NMSV_SEQ_ABSENT 0xFFFFFFFFFFFFFFFFull /* offsets[s]: only the LENGTH of sequence s is supplied. Legal for the sequence of
                                  a read whose mapq test fails while the "prune" option is on: such a read is counted and
                                  never aligned, so the caller need not encode or upload it. */
#define NMSV_MAPQ_NONE (-1)   /* Python None in the seam TSV (count_sread_by_alignment.py:50,288) */

#define NMSV_SLOT_VAR1 0
#define NMSV_SLOT_VAR2 1
#define NMSV_SLOT_REF1 2
#define NMSV_SLOT_REF2 3

/* nmsv_sv.flags */
#define NMSV_SV_SHORT_DEL_DUP 1u /* is_short_del_dup: margin test against ref1 AND ref2 (:312-322,:336-341) */
#define NMSV_SV_SBND 2u          /* single breakend: var1 only, margin against ref1, mapq1 only (:361-397) */

typedef struct nmsv_ctx nmsv_ctx;
typedef struct nmsv_plan nmsv_plan;

typedef struct {
    int32_t match;      /* parasail.matrix_create(alphabet, match, mismatch) */
    int32_t mismatch;   /* <= 0 */
    int32_t gap_open;   /* parasail.ssw(..., open, extend, ...): a gap of length k costs open + (k-1)*extend */
    int32_t gap_extend; /* >= 0 */
} nmsv_scoring;

/* Result fields of parasail.ssw as read at count_sread_by_alignment.py:151-159 (0-based, inclusive).
 * "read" = s1 (the template), "ref" = s2 (the long read).  score1 = -1 (all fields -1): the score is not
 * representable in 16 bits, parasail.ssw returns None (generate_consensus.py:120-125 handles that per pair).
 * There is no limit on sequence lengths: like parasail the engine only gives up when a score actually reaches the
 * 16-bit limit (it reports None for scores above sw_score_cap - 32*match = 32518 with the reference's scoring,
 * parasail above 32764 -- the one documented divergence). */
typedef struct {
    int32_t score1;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
} nmsv_ssw_result;

/* One entry of the dict returned by ssw_check (:162): [score, qstart, qend, tstart, tend, strand], 1-based;
 * q* = template coordinates, t* = read coordinates, strand +1 = '+', -1 = '-', 0 = slot absent / not computed. */
typedef struct {
    int32_t score;
    int32_t qstart;
    int32_t qend;
    int32_t tstart;
    int32_t tend;
    int32_t strand;
} nmsv_aln;

/* One SV candidate (or single-breakend candidate) of a validation batch. The thresholds are the reference's
 * double products turned into integer bounds BY THE CALLER with the reference's own float arithmetic:
 *   score  >  ratio*len   <=>  score  >= score_min   (floor(x)+1)
 *   qstart <  0.1*len     <=>  qstart <= qstart_max  (ceil(x)-1)
 *   qend   >  0.9*len     <=>  qend   >= qend_min    (floor(x)+1)          (:327-333, :379-382)               */
typedef struct {
    uint32_t tmpl[4];       /* sequence index of var1, var2, ref1, ref2; NMSV_NONE = absent. Equal indices share work. */
    uint32_t tmpl_rc[4];    /* NMSV_NONE: the engine derives the reverse complement (A<->T, C<->G, wildcard kept).
                               Otherwise the index of an explicit my_seq.reverse_complement(template) sequence; needed
                               only when the template holds lower-case acgt, which my_seq.py:20-26 leaves
                               uncomplemented while parasail's mapper folds case.                              */
    uint32_t read_begin;    /* reads[read_begin .. read_end) belong to this SV */
    uint32_t read_end;
    int32_t score_min[2];   /* for var1, var2 */
    int32_t qstart_max[2];
    int32_t qend_min[2];
    int32_t margin;         /* var_ref_margin_thres (:171) */
    int32_t min_mapq;       /* var_read_min_mapq */
    uint32_t flags;         /* NMSV_SV_* */
    uint32_t reserved;
} nmsv_sv;

typedef struct {
    uint32_t seq;   /* sequence index of the read (its offsets[] entry may be NMSV_SEQ_ABSENT, see there) */
    uint32_t sv;    /* index of its SV in the svs array */
    int32_t mapq1;  /* NMSV_MAPQ_NONE = None */
    int32_t mapq2;
} nmsv_read;

/* One supporting read: what count_alignment_and_print writes per line of sread_info (:352-358). */
typedef struct {
    uint32_t read;     /* index into the reads array */
    uint32_t sv;
    nmsv_aln aln[4];   /* var1, var2, ref1, ref2 (strand == 0 where the slot is absent) */
} nmsv_record;

/* Work/launch accounting of the last run of a plan (for GCUPS reporting). */
typedef struct {
    uint64_t cells_reference;  /* forward DP cells the reference would compute: sum n*m over every alignment */
    uint64_t cells_executed;   /* forward DP cells computed on the device in the last run (counted on the device) */
    uint64_t cells_padded;     /* cells issued by the statically planned tasks incl. row padding and pipeline skew */
    uint64_t fwd_tasks;        /* statically planned (read, template pair) wavefront tasks */
    uint64_t rev_tasks;        /* begin-coordinate (reversed window) tasks of the last run */
    uint32_t kernel_launches;  /* kernels launched by one nmsv_plan_run */
    uint32_t n_classes;        /* row-strip classes (kernel instantiations) used */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t workspace_bytes;
    /* ABI 2 */
    uint64_t cells_dedup;      /* cells_reference minus the repeats of identical templates of one SV (var1 == var2) */
    uint64_t cells_pruned;     /* cells_dedup - cells_executed: alignments that cannot change any output (reads whose
                                  mapq test fails, :344-346; reference alleles of reads that fail the variant tests, :336-341) */
    uint64_t cells_begin;      /* begin-pass cells of the last run (overhead, in no GCUPS figure) */
    uint64_t fwd_tasks_queued; /* reference-allele forward tasks queued on the device in the last run */
    uint32_t reads_pruned;     /* reads counted as checked without any alignment */
    uint32_t reserved;
} nmsv_stats;

/* ---- context ------------------------------------------------------------------------------------------------ */
int nmsv_abi_version(void);
int nmsv_device_count(void);                          /* usable CUDA devices (0: none) */
int nmsv_create(nmsv_ctx **ctx, int device);          /* one ctx per process/GPU (mirrors --processes, run.py:367) */
void nmsv_destroy(nmsv_ctx *ctx);
const char *nmsv_last_error(const nmsv_ctx *ctx);     /* ctx may be NULL: error of a failed nmsv_create */
/* options: "prune" (default 1): skip alignments that cannot change any output of the stage seam -- reads whose mapq
 * test already fails are only counted (count_sread_by_alignment.py:344-346, :388-389), ref1/ref2 are aligned only for
 * reads that pass the variant tests (:336-341, :384-385). 0: every alignment the reference runs is computed (parity
 * tests; nmsv_plan_download_alns then holds all of them). Counts and records are identical either way. */
int nmsv_set_option(nmsv_ctx *ctx, const char *name, int value);
int nmsv_set_stream(nmsv_ctx *ctx, void *cuda_stream); /* cudaStream_t to launch on; NULL = the ctx's own non-blocking
                                                          stream (pass cudaStreamLegacy to use the default stream) */
int nmsv_device_info(const nmsv_ctx *ctx, int32_t *sm_count, int32_t *sm_clock_khz, int64_t *l2_bytes,
                     int64_t *global_mem_bytes, char *name, size_t name_len);

/* sustained VIADDMNMX.S16x2 issue rate of this GPU in lane-instructions/s (register-only loop): the integer roofline */
int nmsv_measure_dpx_peak(nmsv_ctx *ctx, double *lane_inst_per_s);

/* byte -> code with the mapper of parasail.matrix_create("ACGT", ...): case-insensitive ACGT, rest wildcard */
void nmsv_encode_ascii(const char *ascii, uint64_t n, uint8_t *codes);
/* the same mapper for many sequences that lie inside one ASCII buffer (a block of the seam TSV of
 * count_sread_by_alignment.py:109-111): sequence i = src[src_off[i], +lens[i]) goes to dst[dst_off[i], ...); host only */
void nmsv_encode_pack(const char *src, const uint64_t *src_off, const uint32_t *lens, const uint64_t *dst_off,
                      uint64_t n_seqs, uint8_t *dst);
/* the same with every sequence padded by the wildcard code to the next 16-byte boundary (dst_off[] 16-byte aligned):
 * the destination needs no pre-fill */
void nmsv_encode_pack_padded(const char *src, const uint64_t *src_off, const uint32_t *lens, const uint64_t *dst_off,
                             uint64_t n_seqs, uint8_t *dst);

/* ---- host side of the stage seam: the group-by of count_sread_by_alignment.py:414-448 ------------------------------ */
/* One pass over the key-sorted seam TSV (`key \t read id \t mapq1 \t mapq2 \t sequence`, :111,:120) held in memory
 * (a mapping of the file): consecutive lines with the same key form a group (one SV candidate); inside a group the
 * reference's dict semantics are applied (read id = first blank-separated token, pyssw.py:41; the last record of an id
 * gives its mapq, its sequence is the last non-empty one, an id without any sequence is dropped; "None" = NMSV_MAPQ_NONE;
 * is_sbnd: mapq2 is None). Whole groups are consumed from `start` until at least max_bases read bases are collected (and
 * at least tail_bytes of the file remain: a short rest is not left for a batch of its own) or a buffer is full; *next
 * is where the next call starts. All offsets point into `data`: the caller encodes the sequences
 * straight from there with nmsv_encode_pack. Host only, no ctx. NMSV_E_INVALID: malformed line at *error_off;
 * NMSV_E_CAPACITY: the first group alone does not fit the buffers. */
typedef struct {
    uint64_t key_off;      /* the key column of the group */
    uint32_t key_len;
    uint32_t read_begin;   /* reads[read_begin .. read_end) in first-seen order */
    uint32_t read_end;
    uint32_t reserved;
} nmsv_seam_group;

typedef struct {
    uint64_t id_off;
    uint64_t seq_off;      /* stripped sequence */
    uint32_t id_len;
    uint32_t seq_len;
    int32_t mapq1;
    int32_t mapq2;
} nmsv_seam_read;

int nmsv_seam_scan(const char *data, uint64_t size, uint64_t start, uint64_t max_bases, uint64_t tail_bytes, int is_sbnd,
                   nmsv_seam_group *groups, uint32_t group_cap, nmsv_seam_read *reads, uint32_t read_cap,
                   uint32_t *n_groups, uint32_t *n_reads, uint64_t *next, uint64_t *error_off);

/* ---- operator seam: batched parasail.ssw ---------------------------------------------------------------------- */
int nmsv_ssw_batch(nmsv_ctx *ctx, const uint8_t *codes, uint64_t n_codes, const uint64_t *offsets,
                   const uint32_t *lens, uint32_t n_seqs, const uint32_t *tmpl_idx, const uint32_t *read_idx,
                   uint64_t n_pairs, const nmsv_scoring *scoring, int want_begin, nmsv_ssw_result *out);

/* ---- function seam: batched ssw_check (template and its reverse complement, strand rule, 1-based) ------------- */
int nmsv_ssw_check_batch(nmsv_ctx *ctx, const uint8_t *codes, uint64_t n_codes, const uint64_t *offsets,
                         const uint32_t *lens, uint32_t n_seqs, const uint32_t *tmpl_idx,
                         const uint32_t *tmpl_rc_idx /* may be NULL; entries may be NMSV_NONE */,
                         const uint32_t *read_idx, uint64_t n_pairs, const nmsv_scoring *scoring, nmsv_aln *out);

/* ---- stage seam: supporting-read validation of a batch of SV candidates --------------------------------------- */
/* One-shot, host buffers in / host buffers out (H2D + kernels + D2H inside). checked[i] / supporting[i] are the
 * two trailing columns of the sread_count line of SV i; records (unordered) describe the supporting reads.      */
int nmsv_validate_batch(nmsv_ctx *ctx, const uint8_t *codes, uint64_t n_codes, const uint64_t *offsets,
                        const uint32_t *lens, uint32_t n_seqs, const nmsv_sv *svs, uint32_t n_svs,
                        const nmsv_read *reads, uint32_t n_reads, const nmsv_scoring *scoring,
                        int32_t *checked, int32_t *supporting, nmsv_record *records, uint64_t record_capacity,
                        uint64_t *n_records);

/* The same, split: nmsv_plan_create uploads and plans; nmsv_plan_run launches the whole device step on the
 * ctx stream (asynchronous, inputs resident in HBM, no host synchronisation inside); nmsv_plan_download waits
 * for the stream and copies results out.  A plan may be run any number of times.                              */
int nmsv_plan_create(nmsv_ctx *ctx, nmsv_plan **plan, const uint8_t *codes, uint64_t n_codes,
                     const uint64_t *offsets, const uint32_t *lens, uint32_t n_seqs, const nmsv_sv *svs,
                     uint32_t n_svs, const nmsv_read *reads, uint32_t n_reads, const nmsv_scoring *scoring);
int nmsv_plan_run(nmsv_plan *plan);
int nmsv_plan_download(nmsv_plan *plan, int32_t *checked, int32_t *supporting, nmsv_record *records,
                       uint64_t record_capacity, uint64_t *n_records);
int nmsv_plan_stats(const nmsv_plan *plan, nmsv_stats *stats);
/* measurement hooks: per-launch durations by CUDA events recorded on the launching stream (phase: 0 plan, 1 forward
 * wavefront, 2 decide, 3 follow-up wavefront launch over the dynamic queue, 4 final; rclass = strip height R of a
 * wavefront launch). nmsv_plan_kernel_times returns the MEAN over the runs since nmsv_plan_set_profiling(plan, 1)
 * (the last 32 at most); it waits for those runs, nmsv_plan_run itself never synchronises. */
int nmsv_plan_set_profiling(nmsv_plan *plan, int on);
int nmsv_plan_kernel_times(nmsv_plan *plan, float *ms, uint32_t *phase, uint32_t *rclass, uint32_t capacity,
                           uint32_t *n_launches);
/* copy supporting[n_svs] (int32) of the last run into a caller-owned DEVICE buffer on the ctx stream, e.g. the
 * send buffer of the NCCL gather that merges the per-GPU SV shards */
int nmsv_plan_copy_counts_device(nmsv_plan *plan, void *d_supporting_int32);
/* per (read, slot) alignment tuples of the last run, reads*4 entries; begin coordinates (qstart for '+',
 * qend for '-', tstart) are only filled for reads that reached the begin pass (debug / parity tests) */
int nmsv_plan_download_alns(nmsv_plan *plan, nmsv_aln *alns, uint8_t *read_flags);
void nmsv_plan_destroy(nmsv_plan *plan);

/* ---- "next" row: breakpoint refinement DP, smith_waterman.sw_jump -------------------------------------------------- */
/* Replaces nanomonsv/smith_waterman.py:5-127 (sw_jump), called from nanomonsv/locate_bp.py:37,67: two linear-gap DP
 * matrices (contig x region1, contig x region2) where the second may jump from the best row of the first at a cost of
 * jump_cost * ln(distance), followed by the traceback. Sequences are RAW BYTES here (the reference compares
 * characters, so 'N' matches 'N' and case matters); bytes/offsets/lens use the same one-buffer layout as above.   */
typedef struct {
    int32_t match_score, mismatch_penalty, gap_cost, jump_cost, minimum_score;   /* sw_jump defaults: 1, 2, 2, 2, 10 */
} nmsv_jump_params;

typedef struct {
    int32_t found;                 /* 0: the reference returns None */
    int32_t score;                 /* H2.max() */
    int32_t i1_start, i1_end, i2_start, i2_end;   /* contig coordinates of the two aligned fragments (as returned) */
    int32_t j1_start, j1_end;      /* region1 */
    int32_t j2_start, j2_end;      /* region2 */
    uint32_t n_ops2, n_ops1;       /* traceback operations written to `ops`: H2 part then H1 part, codes 1 = diagonal,
                                      2 = contig base against a gap, 3 = region base against a gap, 4 = the jump cell */
} nmsv_jump_result;

/* ops of problem p start at sum over q < p of (2*n_q + m1_q + m2_q + 4) bytes (n = contig length, m = region
 * lengths); ops_capacity must cover the sum over all problems; only the first n_ops2 + n_ops1 bytes of a problem's
 * segment are written, the rest of it is left as the caller passed it. log_table[d] = ln(d) for d < log_table_len may be
 * supplied by the caller (numpy's log in the Python layer, so that results are bit-identical to the reference);
 * distances beyond the table use the C library's log. */
int nmsv_sw_jump_batch(nmsv_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *offsets,
                       const uint32_t *lens, uint32_t n_seqs, const uint32_t *contig_idx, const uint32_t *region1_idx,
                       const uint32_t *region2_idx, uint32_t n_problems, const nmsv_jump_params *params,
                       const double *log_table, uint32_t log_table_len, nmsv_jump_result *out, uint8_t *ops,
                       uint64_t ops_capacity);

/* Host-side companion (no GPU work, no context): the two alignment rows of sw_jump's return value -- the contig row and the
 * region row, each "<H1 part> <H2 part>" with '-' for gaps (nanomonsv/smith_waterman.py:61-125) -- written out from the
 * traceback operations that nmsv_sw_jump_batch returned. The rows of problem p start at the same offset as its `ops`
 * segment (row_offsets[p] = sum over q < p of 2*n_q + m1_q + m2_q + 4) in `contig_rows` / `region_rows` and are
 * n_ops1 + 1 + n_ops2 bytes long; nothing is written for a problem with found == 0. Returns NMSV_E_INVALID if a
 * traceback walks off its sequences (the operations do not belong to these sequences). */
int nmsv_sw_jump_strings(const uint8_t *bytes, const uint64_t *offsets, const uint32_t *lens, const uint32_t *contig_idx,
                         const uint32_t *region1_idx, const uint32_t *region2_idx, uint32_t n_problems,
                         const nmsv_jump_result *results, const uint8_t *ops, const uint64_t *row_offsets,
                         uint8_t *contig_rows, uint8_t *region_rows);

#ifdef __cplusplus
}
#endif
#endif /* NMSV_H */
