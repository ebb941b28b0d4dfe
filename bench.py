#!/usr/bin/env python3
"""bench.py -- supporting-read validation throughput (GCUPS, SV candidates/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # host-CPU "parasail-equivalent" port

Headline workload = BASELINE.json configs[1] ("synthetic COLO829-shaped": canonical SV candidates, 30x tumor + 30x control
ONT-like reads ~10 kb at ~8 % error, +-1 kb templates -> m = 2000). One STEP validates one batch of --svs-per-gpu
candidates (tumor AND control reads) per GPU; the 20k-candidate job of the config is that step repeated, so GCUPS and
candidates/s do not depend on the batch count. Weak scaling: every rank gets its own batch (own seed, same planned work),
no data-path collective; for N > 1 the per-SV counts are merged with one NCCL all-reduce at the end of the job.

value  = reference-equivalent forward DP cells (sum 2*m*n over every alignment parasail would run) / device time,
         inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e    = the same cells / wall time of nmsv_validate_batch called with pinned HOST buffers (H2D, planning, kernels, D2H
         inside the timed region), measured two ways every run: two calls in flight (two contexts, two host threads, so that
         the upload of one step overlaps the kernels of the previous one: `e2e.pipelined`) and one call at a time
         (`e2e.serial`); `e2e.value` is the faster of the two and `calls_in_flight` says which.
The device does not compute every cell the reference computes: identical templates of one SV (var1 == var2 without an
insertion) are aligned once, reads whose mapping-quality test already fails are only counted, and reference-allele
alignments are run only for reads that pass the variant tests -- the outputs are identical (`parity_sampled`, tests/).
`cells_executed` (counted on the device) is what the roofline fraction and `gcups_executed` use; `cells_pruned` says how
much was skipped.

The same line carries, under `configs`, one entry per BASELINE.json config (1, 3, 4 as single-GPU shape measurements at
N = 1; 5 = the 10^6-pair kernel sweep, sharded over the N ranks) and, under `parity_sampled`, a comparison of the CUDA
results with the CPU oracle on a sample of every config (the run aborts non-zero on a mismatch); for N > 1 also a
`strong` section: ONE fixed job partitioned over the ranks with the product's own sharding (nanomonsv_b200/shard.py).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# DPX lane-instructions per DP cell on the hot path of sw_wavefront_kernel<16,3,1>: per step and lane 16 packed
# (template, revcomp) cell pairs cost 48 VIADDMNMX.S16x2 + 16 VIMNMX3.S16x2 (x, H, E', F') + 2 VIADDMNMX.S16x2 for
# the sampled running maximum = 66 instructions per 32 cells (SASS count)
DPX_PER_CELL = 66.0 / 32.0
WORKLOAD = "colo829_shaped"
# mean statically PLANNED forward cells of one synthetic candidate of the headline workload (tumor + control reads that
# pass the mapq test, both strands, identical variant templates counted once; the lazily queued reference-allele
# alignments come on top). The per-GPU batch is cut at --svs-per-gpu times this budget: equal device work per rank.
CELLS_PER_CANDIDATE = 3.1e9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout. Native libraries (NCCL's version banner, ...) write to fd 1 directly, so
# fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ----------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------
def _mapq_ok(rd, is_sbnd: bool, min_mapq: int) -> bool:
    if rd.mapq1 is None or rd.mapq1 < min_mapq:
        return False
    return is_sbnd or (rd.mapq2 is not None and rd.mapq2 >= min_mapq)


def build_batch(n_sv: int, rank: int, args=None, preset: str = WORKLOAD, budget: float | None = -1.0, **over):
    """One step's batch for one GPU: tumor reads then control reads of the same candidates (canonical or single
    breakend). budget = planned cells to cut the batch at (-1: CELLS_PER_CANDIDATE * n_sv for the headline workload,
    None: keep all n_sv candidates)."""
    from nanomonsv_b200 import synth
    from nanomonsv_b200.batch import BatchBuilder
    from nanomonsv_b200.count_sread_by_alignment import canonical_templates, sbnd_templates
    base = synth.PRESETS[preset]
    kw = dict(seed=base.seed + 7919 * rank, contig_len=4_000_000, n_contigs=4, stratified=True)
    kw.update(over)
    cfg = dataclasses.replace(base, **kw)
    t0 = time.time()
    if budget is not None and budget < 0:
        budget = CELLS_PER_CANDIDATE * n_sv
    spare = max(8, n_sv // 5) if budget is not None else 0
    wl = synth.make_workload(cfg, n_sv + spare)
    L = cfg.validate_sequence_length
    tmpl = {}
    for i, sv in enumerate(wl.svs):
        if sv.is_sbnd:
            v1, r1 = sbnd_templates(sv.key, wl.fasta, L)
            tmpl[i] = ([v1, None, r1, None], True, False)
        else:
            v1, v2, r1, r2, short = canonical_templates(sv.key, wl.fasta, L)
            tmpl[i] = ([v1, v2, r1 if short else None, r2 if short else None], False, short)
    keep = list(range(len(wl.svs)))
    if budget is not None:
        # weak scaling needs equal work per rank: keep the first candidates (in generation order) whose planned cell
        # count reaches the per-GPU budget
        keep, acc = [], 0.0
        for i in range(len(wl.svs)):
            if acc >= budget:
                break
            t, is_sbnd, short = tmpl[i]
            rows = sum(len(x) for x in {t[0], t[1]} if x)
            bases = sum(len(r.codes) for smp in ("tumor", "control") for r in wl.samples[smp][i] if _mapq_ok(r, is_sbnd, 0))
            # + the reference alleles that the device aligns lazily: nearly every read across a short deletion / duplication
            # passes the variant tests and has its first reference allele aligned until it comes within the margin
            # (measured: ~0.6 of the read on average)
            ref_rows = 0.6 * len(t[2]) if (short and t[2]) else 0.0
            keep.append(i); acc += 2.0 * (rows + ref_rows) * bases
    keep_set = set(keep)
    bb = BatchBuilder(score_ratio_thres=1.2, var_read_min_mapq=0)
    for sample in ("tumor", "control"):
        for i in [j for j in wl.sv_order() if j in keep_set]:
            t, is_sbnd, short = tmpl[i]
            bb.add_sv(t, is_sbnd, short)
            for rd in wl.samples[sample][i]:
                bb.add_read(rd.codes, rd.mapq1, None if is_sbnd else rd.mapq2)
    pack, svs, reads = bb.finalize()
    build_batch.last = (wl, [j for j in wl.sv_order() if j in keep_set])      # for the stage-seam measurement
    log(f"[rank {rank}] {preset}: {len(keep)} candidates x (tumor, control), {len(reads)} reads, "
        f"{pack.finalize()[0].size / 1e6:.0f} MB of bases, built in {time.time() - t0:.1f}s")
    return cfg, pack, svs, reads


# ----------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle on a bounded sample (timed = cpu_baseline / reference arm; results = parity check)
# ----------------------------------------------------------------------------------------------------------------
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


class OracleJob:
    """Every alignment the reference would run for the first SVs of a packed batch: each present template and its
    reverse complement against every read of the SV (no sharing of identical templates, like the reference)."""

    def __init__(self, pack, svs, reads, target_cells: float):
        from nanomonsv_b200 import _lib
        codes, offsets, lens = pack.finalize()
        extra, rc_of = [], {}
        pos = int(codes.size)
        off_l, len_l = [], []
        t_idx, r_idx, rows, slots = [], [], [], []
        cells = 0
        n_used = 0
        for i, sv in enumerate(svs):
            if cells >= target_cells:
                break
            sbnd = bool(sv["flags"] & _lib.SV_SBND)
            rb, re_ = int(sv["read_begin"]), int(sv["read_end"])
            seqs = reads["seq"][rb:re_].astype(np.int64)
            for slot in range(4):
                t = int(sv["tmpl"][slot])
                if t == _lib.NMSV_NONE or (sbnd and slot in (1, 3)):
                    continue
                rc = int(sv["tmpl_rc"][slot])
                if rc == _lib.NMSV_NONE:
                    rc = rc_of.get(t)
                    if rc is None:
                        tc = codes[int(offsets[t]):int(offsets[t]) + int(lens[t])]
                        rc = len(lens) + len(off_l)
                        rc_of[t] = rc
                        extra.append(_COMP[tc[::-1]]); off_l.append(pos); len_l.append(len(tc)); pos += len(tc)
                n = len(seqs)
                t_idx.append(np.stack([np.full(n, t), np.full(n, rc)], axis=1).ravel())
                r_idx.append(np.repeat(seqs, 2))
                rows.append(np.arange(rb, re_)); slots.append(np.full(n, slot))
                cells += 2 * int(lens[t]) * int(lens[seqs].sum())
            n_used += 1
        self.n_svs, self.n_reads = n_used, int(svs["read_end"][n_used - 1]) if n_used else 0
        self.cells = cells
        self.codes = np.concatenate([codes] + extra) if extra else codes
        self.offsets = np.concatenate([offsets, np.asarray(off_l, dtype=np.uint64)])
        self.lens = np.concatenate([lens, np.asarray(len_l, dtype=np.uint32)])
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)   # noqa: E731
        self.t_idx, self.r_idx = cat(t_idx, np.uint32), cat(r_idx, np.uint32)
        self.rows, self.slots = cat(rows, np.int64), cat(slots, np.int64)
        self.m = self.lens[self.t_idx[0::2]].astype(np.int64) if len(self.t_idx) else np.zeros(0, dtype=np.int64)

    def run(self, threads: int):
        """(seconds, threads used, results) of the striped AVX2 port with the reversed begin pass for every alignment."""
        from oracle import oracle as O
        t0 = time.perf_counter()
        res, used = O.ssw_batch(self.codes, self.offsets, self.lens, self.t_idx, self.r_idx, engine="striped", threads=threads)
        return time.perf_counter() - t0, used, res

    def expected(self, res):
        """ssw_check's 6-tuple per (read, slot) from the oracle's two strands (count_sread_by_alignment.py:151-160)."""
        f, r = res[0::2], res[1::2]
        plus = f["score1"] > r["score1"]                       # ties -> '-'
        m = self.m
        out = np.zeros(len(f), dtype=[("score", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("tstart", "<i4"),
                                      ("tend", "<i4"), ("strand", "<i4")])
        out["score"] = np.where(plus, f["score1"], r["score1"])
        out["qstart"] = np.where(plus, f["read_begin1"] + 1, m - r["read_end1"])
        out["qend"] = np.where(plus, f["read_end1"] + 1, m - r["read_begin1"])
        out["tstart"] = np.where(plus, f["ref_begin1"] + 1, r["ref_begin1"] + 1)
        out["tend"] = np.where(plus, f["ref_end1"] + 1, r["ref_end1"] + 1)
        out["strand"] = np.where(plus, 1, -1)
        return out


def parity_validate(eng, pack, svs, reads, job: OracleJob, res, full_counts=None) -> dict:
    """CUDA path vs the oracle on the sampled SVs: a plan that computes EVERY alignment (prune off) must reproduce score,
    strand and end cell of every (read, template) and the whole 6-tuple wherever the begin pass ran; the default
    (pruning) plan must give the same counts and records. Returns {checked, equal}; raises on a mismatch."""
    n_sv, n_rd = job.n_svs, job.n_reads
    exp = job.expected(res)
    plan = eng.plan(pack, svs[:n_sv], reads[:n_rd], prune=False)
    plan.run()
    c_full, s_full, r_full = plan.download()
    alns, flags = plan.download_alns()
    plan.close()
    got = alns[job.rows, job.slots]
    cand = (flags[job.rows] & 1) != 0
    plus = exp["strand"] > 0
    ok = (got["score"] == exp["score"]) & (got["strand"] == exp["strand"]) & (got["tend"] == exp["tend"])
    ok &= np.where(plus, got["qend"] == exp["qend"], got["qstart"] == exp["qstart"])          # end cell
    full = (got["qstart"] == exp["qstart"]) & (got["qend"] == exp["qend"]) & (got["tstart"] == exp["tstart"])
    ok &= np.where(cand & (exp["score"] > 0), full, True)
    checked, equal = int(len(exp)), int(ok.sum())
    plan = eng.plan(pack, svs[:n_sv], reads[:n_rd])
    plan.run()
    c_p, s_p, r_p = plan.download()
    plan.close()
    same = bool((c_p == c_full).all() and (s_p == s_full).all() and (r_p == r_full).all())
    if full_counts is not None:      # the sampled SVs inside the full batch
        same = same and bool((full_counts[:n_sv] == s_p).all())
    out = {"checked": checked, "equal": equal, "begin_pass_tuples": int((cand & (exp["score"] > 0)).sum()),
           "sv_candidates": n_sv, "supporting_reads": int(s_full.sum()), "pruned_plan_identical": same}
    if equal != checked or not same:
        bad = np.flatnonzero(~ok)[:5]
        raise SystemExit(f"PARITY MISMATCH against the oracle: {out}; first rows {[(int(job.rows[b]), int(job.slots[b])) for b in bad]}")
    return out


def stage_seam_e2e(eng, wl, order, n_cand, checked, supporting, n_records) -> dict:
    """The drop-in stage function itself (count_sread_from_seam = count_sread_by_alignment after the read gathering,
    reference :414-448) on seam TSVs written before the timer: parsing, group-by, template construction, packing,
    H2D, kernels, D2H and the two output files per sample, tumor then control like `nanomonsv get`. The files are held
    to the device-resident results of the same candidates."""
    import shutil
    import tempfile
    from nanomonsv_b200 import count_sread_by_alignment as C
    from nanomonsv_b200 import synth
    tmp = tempfile.mkdtemp(prefix="nmsv_stage_")
    try:
        seams, size = {}, 0
        for sample in ("tumor", "control"):
            seams[sample] = os.path.join(tmp, sample + ".seam.tsv")
            with open(seams[sample], "w") as h:
                for i in order:
                    key = wl.svs[i].key
                    for rd in wl.samples[sample][i]:
                        h.write(f"{key}\t{rd.rid}\t{rd.mapq1}\t{rd.mapq2}\t{synth.codes_to_str(rd.codes)}\n")
            size += os.path.getsize(seams[sample])
        L = wl.config.validate_sequence_length

        def run():
            t0 = time.perf_counter()
            for sample in ("tumor", "control"):
                C.count_sread_from_seam(seams[sample], seams[sample] + ".count", seams[sample] + ".info", wl.fasta, False, 0,
                                        1.2, engine=eng, validate_sequence_length=L)
            return time.perf_counter() - t0
        run()                       # warm-up (page cache, pools)
        dt = min(run(), run())
        sweep = {}
        if os.environ.get("NMSV_BENCH_STAGE_SWEEP"):      # batch-size sweep (measurement knob)
            default = C.BATCH_BASES
            for fl in (2, 3, 4):
                for mb in (16, 32, 64):
                    C.BATCH_BASES, C.INFLIGHT = mb << 20, fl
                    run()
                    sweep[f"{mb}MB x{fl}"] = n_cand / min(run(), run())
            C.INFLIGHT = 4
            C.BATCH_BASES = default
            log("[stage seam] candidates/s by batch size:", {k: round(v) for k, v in sweep.items()})
        got_c, got_s, n_info = [], [], 0
        for sample in ("tumor", "control"):
            for line in open(seams[sample] + ".count"):
                f = line.rstrip("\n").split("\t")
                got_c.append(int(f[-2])); got_s.append(int(f[-1]))
            n_info += sum(1 for _ in open(seams[sample] + ".info"))
        same = got_c == [int(x) for x in checked] and got_s == [int(x) for x in supporting] and n_info == n_records
        if not same:
            raise SystemExit("stage seam: the files written by count_sread_from_seam differ from the device-resident results")
        # A job of the size the reference runs (thousands of candidates per sample) spends its time in the steady state of the
        # pipeline; the two short calls above pay its fill and drain twice per 0.2 s. The same candidates REPS times over under
        # different SV ids (the last field of the key; the templates only depend on the others):
        REPS = max(1, min(4, 1300 // max(1, n_cand)))      # ~1 200 candidates per sample
        long_job = None
        if REPS > 1:
            for sample in ("tumor", "control"):
                tails = [(wl.svs[i].key, [f"\t{rd.rid}\t{rd.mapq1}\t{rd.mapq2}\t{synth.codes_to_str(rd.codes)}\n"
                                          for rd in wl.samples[sample][i]]) for i in order]
                with open(seams[sample] + ".long", "w") as h:
                    for rep in range(REPS):
                        for key, lines in tails:
                            k = key + (f"r{rep}" if rep else "")
                            h.write("".join(k + t for t in lines))
                del tails

            def run_long():
                t0 = time.perf_counter()
                for sample in ("tumor", "control"):
                    C.count_sread_from_seam(seams[sample] + ".long", seams[sample] + ".lcount", seams[sample] + ".linfo", wl.fasta, False, 0,
                                            1.2, engine=eng, validate_sequence_length=L)
                return time.perf_counter() - t0
            dt_long = min(run_long(), run_long())
            lc, ls, l_info = [], [], 0
            for sample in ("tumor", "control"):
                rows = [line.rstrip("\n").split("\t") for line in open(seams[sample] + ".lcount")]
                lc += [[int(f[-2]) for f in rows[r * len(order):(r + 1) * len(order)]] for r in range(REPS)]
                ls += [[int(f[-1]) for f in rows[r * len(order):(r + 1) * len(order)]] for r in range(REPS)]
                l_info += sum(1 for _ in open(seams[sample] + ".linfo"))
            half = len(order)
            exp_c = [[int(x) for x in checked[:half]]] * REPS + [[int(x) for x in checked[half:]]] * REPS
            exp_s = [[int(x) for x in supporting[:half]]] * REPS + [[int(x) for x in supporting[half:]]] * REPS
            if lc != exp_c or ls != exp_s or l_info != REPS * n_records:
                raise SystemExit("stage seam (long job): the files differ from the device-resident results")
            long_job = {"sv_per_s": REPS * n_cand / dt_long, "s": dt_long, "candidates_per_sample": REPS * n_cand,
                        "seam_gb_per_s": REPS * size / dt_long / 1e9, "files_equal_device_resident_results": True,
                        "what": f"the same two calls on seam files holding the candidates {REPS} times over (different SV ids): "
                                "the fill and drain of the pipeline weigh less, as in a job of the reference's size"}
        return {"sv_per_s": n_cand / dt, "s": dt, "seam_bytes": size, "seam_gb_per_s": size / dt / 1e9,
                "files_equal_device_resident_results": same,
                "long_job": long_job,
                "what": "count_sread_from_seam(tumor) + (control): seam TSV on disk (page cache) -> sread_count / sread_info files; "
                        "one Python process, one GPU; nmsv_seam_scan + nmsv_encode_pack on the host, GPU batches on a worker thread"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception as exc:   # noqa: BLE001
            log("nvidia-smi unavailable:", exc)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:   # noqa: BLE001
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (parasail is not installable here -> the oracle port)
# ----------------------------------------------------------------------------------------------------------------
def arm_config(n_cand: int, n_reads: int, code_bytes: int, l2_bytes: int, world: int) -> dict:
    """`config` of the JSON line, the same object in both arms: the workload and the batch one GPU takes per step (rank 0's
    batch; the CPU arm times a bounded sample of that batch per step and says which in cpu_baseline.sample)."""
    return {"workload": f"{WORKLOAD} (BASELINE.json configs[1]): canonical SV candidates, 30x tumor + 30x control "
                        f"ONT-like reads ~10 kb, 8% error, m=2000 templates",
            "sv_candidates_per_gpu_per_step": int(n_cand), "reads_per_gpu_per_step": int(n_reads),
            "read_bases_per_gpu_mb": round(code_bytes / 1e6, 1),
            "l2": f"inputs ({code_bytes / 1e6:.0f} MB of read bases per GPU) exceed the {l2_bytes / 1e6:.0f} MB L2",
            "parallelism": f"sv-shards x{world}, equal planned cells per GPU, one NCCL all-reduce of the count tables at the end of the job" if world > 1 else "single GPU"}


def reference_arm(args, rank: int):
    if rank != 0:
        return
    # rank 0's batch of the other arm (same generator, same seed, same size): each step times a bounded sample of it
    cfg, pack, svs, reads = build_batch(args.svs_per_gpu, 0, args)
    threads = os.cpu_count() or 1
    l2_bytes = 132644864          # B200; read from the device when there is one
    try:
        import torch
        if torch.cuda.is_available():
            l2_bytes = int(torch.cuda.get_device_properties(0).L2_cache_size)
    except Exception:
        pass
    # calibrate on a small sample, then size one step to ~args.cpu_seconds
    job = OracleJob(pack, svs, reads, 2e10)
    dt, used, _ = job.run(threads)
    job = OracleJob(pack, svs, reads, max(2e10, job.cells / dt * args.ref_step_seconds))
    for _ in range(args.warmup):
        job.run(threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, used, _ = job.run(threads)
    dt = (time.perf_counter() - t0) / args.steps
    gcups = job.cells / dt / 1e9
    sample = (f"{job.n_svs} of the batch's SV candidates (tumor reads), {len(job.t_idx)} alignments incl. reverse-complement "
              f"strand and reversed begin pass, {job.cells / 1e9:.1f} G forward cells per step")
    line = {"impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8->int16 striped SIMD (AVX2)", "data": "synthetic",
            "config": arm_config(len(svs) // 2, len(reads), pack.finalize()[0].size, l2_bytes, args.gpus),
            "sv_per_s": job.n_svs / 2 / dt,
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": used, "kind": "port", "sample": sample,
                             "engine": "oracle/nmsv_striped.c (parasail-equivalent port; parasail itself is not installable here)"},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# secondary configs (BASELINE.json configs 1, 3, 4: shapes; 5: the pair sweep)
# ----------------------------------------------------------------------------------------------------------------
SHAPES = [
    # (BASELINE config number, preset, candidates, overrides)
    (1, "testdata_shaped", 400, dict(contig_len=2_000_000, n_contigs=2)),
    (3, "insertion_heavy", 120, dict(contig_len=2_000_000, n_contigs=2)),
    (4, "single_breakend", 2000, dict(contig_len=2_000_000, n_contigs=2)),
]


def time_runs(torch, stream, fn, warmup: int, steps: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def shape_entry(torch, stream, eng, dpx_peak, number, preset, n_sv, over, cpu_gcups, cpu_seconds, threads):
    cfg, pack, svs, reads = build_batch(n_sv, 0, None, preset=preset, budget=None, **over)
    plan = eng.plan(pack, svs, reads)
    plan.set_profiling(True)
    ms = time_runs(torch, stream, plan.run, 2, 3)
    kt = plan.kernel_times()
    _, supporting, _ = plan.download()
    st = plan.stats()
    plan.close()
    fwd = [(r, t) for name, r, t in kt if name == "sw_forward"]
    fwd_ms = sum(t for _, t in fwd)
    entry = {"baseline_config": number, "workload": preset, "sv_candidates": len(svs) // 2, "reads": int(len(reads)),
             "template_rows": 2 * cfg.validate_sequence_length, "ms": ms, "sv_per_s": (len(svs) // 2) / (ms * 1e-3),
             "gcups": st["cells_reference"] / ms / 1e6, "gcups_executed": st["cells_executed"] / ms / 1e6,
             "cells_reference": st["cells_reference"], "cells_executed": st["cells_executed"],
             "cells_pruned": st["cells_pruned"], "supporting_reads": int(supporting.sum()),
             "roofline": {"frac": st["cells_executed"] * DPX_PER_CELL / (fwd_ms * 1e-3) / dpx_peak,
                          "kernels": [f"sw_wavefront_kernel<{r}>" for r, _ in fwd], "kernel_ms": fwd_ms,
                          "note": "all forward launches of the step together, DPX cost of the R=16 instantiation"}}
    job = OracleJob(pack, svs, reads, cpu_gcups * 1e9 * cpu_seconds)
    _, _, res = job.run(threads)
    entry["parity_sampled"] = parity_validate(eng, pack, svs, reads, job, res, supporting)
    log(f"[config {number}] {preset}: {entry['gcups_executed']:.0f} GCUPS executed, {entry['sv_per_s']:.0f} SV/s, parity "
        f"{entry['parity_sampled']['equal']}/{entry['parity_sampled']['checked']}")
    return entry


def pair_sweep_entry(torch, dist, dev, eng, rank, world, total_pairs, cpu_gcups, cpu_seconds, threads):
    """BASELINE config 5: 10^6 (read, template) pairs, m = 400, read length uniform in 1-50 kb, drawn from a pool of
    reads, streamed through nmsv_ssw_check_batch (both strands, begin pass, host buffers) in batches; the pairs are
    dealt to the ranks, no collective on the data path."""
    from nanomonsv_b200 import SeqPack, synth
    from oracle import oracle as O
    t0 = time.time()
    seqs, t_idx, r_idx = synth.make_pair_sweep(total_pairs, m=400, n_min=1000, n_max=50000, pool_reads=4000, seed=20261019 + 5)
    pack = SeqPack()
    for s in seqs:
        pack.add_codes(s)
    codes, offsets, lens = pack.finalize()
    mine = slice(rank, total_pairs, world)
    t_my, r_my = t_idx[mine], r_idx[mine]
    batch = 65536
    chunks = [(t_my[a:a + batch], r_my[a:a + batch]) for a in range(0, len(t_my), batch)]
    cells_my = float((2 * 400 * lens[r_my].astype(np.int64)).sum())
    log(f"[rank {rank}] pair sweep: {len(t_my)} of {total_pairs} pairs in {len(chunks)} batches, pool of {len(seqs)} sequences "
        f"({codes.size / 1e6:.0f} MB), built in {time.time() - t0:.1f}s")
    eng.ssw_check_batch(pack, chunks[0][0][:4096], chunks[0][1][:4096])      # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = [eng.ssw_check_batch(pack, t, r) for t, r in chunks]
    torch.cuda.synchronize()
    dt_serial = time.perf_counter() - t0
    # the same batches with two calls in flight (two contexts, two host threads; ctypes releases the GIL), the way a caller
    # with a stream of batches drives the library: the upload of a batch and the end of its launch (the longest reads of a
    # batch finish last) run under the kernels of the other one
    from concurrent.futures import ThreadPoolExecutor
    from nanomonsv_b200 import Engine
    eng2 = Engine(dev.index)
    engines = (eng, eng2)
    eng2.ssw_check_batch(pack, chunks[0][0][:4096], chunks[0][1][:4096])     # warm-up of the second context
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ThreadPoolExecutor(max_workers=2) as pool:
        t0 = time.perf_counter()
        outs2 = list(pool.map(lambda kc: engines[kc[0] & 1].ssw_check_batch(pack, kc[1][0], kc[1][1]), enumerate(chunks)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    eng2.close()
    if any((a != b).any() for a, b in zip(outs, outs2)):
        raise SystemExit("pair sweep: the results of the calls in flight differ from the serial ones")
    red = torch.tensor([dt, dt_serial], dtype=torch.float64, device=dev)
    tot = torch.tensor([cells_my, float(len(t_my))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    dt, dt_serial = float(red[0]), float(red[1])
    out = np.concatenate(outs)
    entry = {"baseline_config": 5, "workload": "pair_sweep: m=400, reads 1-50 kb uniform, half carry the template at 8% error",
             "pairs": int(tot[1]), "n_gpus": world, "s": dt, "gcups": float(tot[0]) / dt / 1e9, "pairs_per_s": float(tot[1]) / dt,
             "hits": int((out["score"] > 480).sum()), "calls_in_flight": 2,
             "serial": {"s": dt_serial, "gcups": float(tot[0]) / dt_serial / 1e9, "note": "one call at a time"},
             "note": "through nmsv_ssw_check_batch (both strands + begin pass) with host buffers: H2D of the pool per batch and "
                     "D2H of the tuples are inside the time; max over ranks"}
    if rank == 0:      # oracle on a sample of this rank's pairs
        n = int(max(64, min(len(t_my), cpu_gcups * 1e9 * cpu_seconds / (2 * 400 * float(lens[r_my].mean())))))
        n_t = int(lens.size)
        tm = np.unique(t_my[:n])
        rc_off, pos = [], int(codes.size)
        extra = []
        for t in tm:
            tc = codes[int(offsets[t]):int(offsets[t]) + int(lens[t])]
            extra.append(_COMP[tc[::-1]]); rc_off.append(pos); pos += len(tc)
        rc_of = {int(t): n_t + k for k, t in enumerate(tm)}
        c2 = np.concatenate([codes] + extra)
        o2 = np.concatenate([offsets, np.asarray(rc_off, dtype=np.uint64)])
        l2 = np.concatenate([lens, lens[tm]])
        ti = np.stack([t_my[:n], np.array([rc_of[int(t)] for t in t_my[:n]], dtype=np.uint32)], axis=1).ravel()
        ri = np.repeat(r_my[:n], 2)
        res, _ = O.ssw_batch(c2, o2, l2, ti, ri, engine="striped", threads=threads)
        job = OracleJob.__new__(OracleJob)
        job.m = np.full(n, 400, dtype=np.int64)
        exp = job.expected(res)
        got = out[:n]
        ok = np.ones(n, dtype=bool)
        for name in ("score", "qstart", "qend", "tstart", "tend", "strand"):
            ok &= (got[name] == exp[name]) | ((exp["score"] == 0) & (name != "score"))
        entry["parity_sampled"] = {"checked": n, "equal": int(ok.sum())}
        if int(ok.sum()) != n:
            raise SystemExit(f"PARITY MISMATCH against the oracle on the pair sweep: {entry['parity_sampled']}")
        log(f"[config 5] pair sweep: {entry['gcups']:.0f} GCUPS on {world} GPU(s), parity {int(ok.sum())}/{n}")
    return entry


# ----------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--svs-per-gpu", type=int, default=int(os.environ.get("NMSV_BENCH_SVS", "1024")),
                    help="candidates drawn per GPU and step (the batch is cut at this many times the planned-cell budget per candidate); a launch of "
                         "~0.45 s: the end of a launch, when the longest reads run at low occupancy, weighs 1 % instead of 3 % at 384")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU work per step of the baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=2.0, help="--impl reference: CPU work per step (a bounded sample of the batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU-oracle leg (cpu_baseline and parity_sampled)")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only")
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs of the config-5 sweep (whole job, all ranks)")
    ap.add_argument("--strong-svs", type=int, default=768, help="candidates of the fixed strong-scaling job (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from nanomonsv_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the validation engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    eng = Engine(local_rank)
    stream = torch.cuda.Stream(device=dev)      # kernels, NCCL and the timing events all live on this stream
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    info = eng.device_info()
    dpx_peak = eng.measure_dpx_peak()

    cfg, pack, svs, reads = build_batch(args.svs_per_gpu, rank, args)
    head_wl, head_order = build_batch.last
    codes, offsets, lens = pack.finalize()
    n_cand = len(svs) // 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident hot path
    plan = eng.plan(pack, svs, reads)
    st = plan.stats()
    # the path's only exchange: ONE merge of the per-GPU supporting-read count tables at the end of the job
    # (NCCL all-reduce of a dense int32 table over NVLink); the job here = the timed steps
    n_max = torch.tensor([len(svs)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
    cap = int(n_max.item())
    counts = torch.zeros((args.steps, max(world, 1), cap), dtype=torch.int32, device=dev)

    def step(k):
        plan.run()
        if world > 1:
            plan.copy_counts_to_device(counts[k, rank].data_ptr())

    def final_gather():
        if world > 1:
            dist.all_reduce(counts)

    for _ in range(args.warmup):
        step(0)
    final_gather()
    counts.zero_()
    barrier()
    plan.set_profiling(True)                          # per-launch events of every timed step (no host sync between them)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step(k)
    final_gather()
    e1.record(stream)
    barrier()
    ms_step = e0.elapsed_time(e1) / args.steps
    ktimes = plan.kernel_times()                      # mean over the timed steps
    plan.set_profiling(False)
    checked, supporting, records = plan.download()
    st = plan.stats()

    # ---- end to end through the C ABI with pinned host buffers
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
    from nanomonsv_b200.engine import SeqPack
    hpack = SeqPack.from_arrays(pin(codes), pin(offsets), pin(lens))
    hsvs, hreads = pin(svs.view(np.uint8)).view(svs.dtype), pin(reads.view(np.uint8)).view(reads.dtype)

    def e2e_step(k):
        c, s, r = eng.validate(hpack, hsvs, hreads)
        if world > 1:
            counts[k, rank, :len(s)] = torch.from_numpy(np.ascontiguousarray(s)).to(dev, non_blocking=True)
        return s, r

    counts.zero_()
    for _ in range(2):
        e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        s_e2e, r_e2e = e2e_step(k)
    barrier()
    e2e_serial_s = (time.perf_counter() - t0) / args.steps        # one call at a time: copies and compute in series

    # The same calls with two in flight, the way a caller with a stream of batches drives the library (the stage function
    # keeps one batch on the GPU while it prepares the next): two contexts, two host threads (ctypes releases the GIL),
    # so the upload of step k+1 runs under the kernels of step k. Every step still uploads its own inputs from pinned
    # host memory and downloads its own results.
    from concurrent.futures import ThreadPoolExecutor
    eng2 = Engine(local_rank)
    engines = (eng, eng2)

    def e2e_call(k):
        return engines[k & 1].validate(hpack, hsvs, hreads)

    with ThreadPoolExecutor(max_workers=2) as pool:
        list(pool.map(e2e_call, range(2)))                         # warm the second context
        counts.zero_()
        barrier()
        t0 = time.perf_counter()
        results = list(pool.map(e2e_call, range(args.steps)))
        for k, (c, s_, r_) in enumerate(results):
            if world > 1:
                counts[k, rank, :len(s_)] = torch.from_numpy(np.ascontiguousarray(s_)).to(dev, non_blocking=True)
        final_gather()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
    assert all((s_ == supporting).all() and (r_ == records).all() for _, s_, r_ in results), "pipelined e2e results differ"
    clocks = sampler.stop()
    assert (s_e2e == supporting).all() and (r_e2e == records).all(), "e2e and device-resident results differ"
    eng2.close()

    # ---- max over ranks, totals over ranks
    red = torch.tensor([ms_step, e2e_s, e2e_serial_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([st["cells_reference"], st["cells_executed"], n_cand, int(supporting.sum()), st["cells_pruned"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_step, e2e_s, e2e_serial_s = float(red[0]), float(red[1]), float(red[2])
    cells_ref, cells_exec, cand_total, supp_total, cells_pruned = [float(x) for x in tot]

    line = None
    threads = os.cpu_count() or 1
    if rank == 0:
        fwd = [(r, ms) for name, r, ms in ktimes if name == "sw_forward"]
        dom_r, dom_ms = max(fwd, key=lambda x: x[1])
        step_kernel_ms = sum(ms for _, _, ms in ktimes)
        achieved = st["cells_executed"] * DPX_PER_CELL / (dom_ms * 1e-3)     # forward cells all run in this launch
        h2d = st["h2d_bytes"]
        traffic, traffic_src = _traffic(st["cells_executed"])
        roofline = {
            "bound": "int_alu (DPX issue rate; not hbm, not tensor)",
            "kernel": f"sw_wavefront_kernel<{dom_r}> (forward launch; it also runs the queued reference-allele alignments and "
                      f"the begin pass, whose cells are not counted)",
            "achieved": achieved / 1e12, "peak": dpx_peak / 1e12, "unit": "T DPX lane-inst/s",
            "frac": achieved / dpx_peak,
            "peak_source": "measured live: register-only VIADDMNMX.S16x2 loop on this GPU (nmsv_measure_dpx_peak); the ALU "
                           "pipe issues one warp instruction per 2 clocks per scheduler",
            "peak_nominal_int32": {"value": info["sm_count"] * 128 * info["sm_clock_khz"] * 1e3 / 1e12, "unit": "T lane-ops/s",
                                   "note": "SMs x 128 INT32 lanes x clock (SURVEY 8d); DPX s16x2 is a half-rate pipe: 64 lanes/clk/SM"},
            "dpx_lane_inst_per_cell": DPX_PER_CELL,
            "kernel_ms": dom_ms, "kernel_ms_source": f"mean over the {args.steps} timed steps (CUDA events on the launching stream)",
            "kernel_share_of_step": dom_ms / step_kernel_ms,
            "gcups_kernel": st["cells_executed"] / (dom_ms * 1e-3) / 1e9,
            "traffic": traffic, "traffic_source": traffic_src,
            "hbm": {"algorithmic_bytes": int(codes.size), "achieved_gbs": codes.size / (dom_ms * 1e-3) / 1e9,
                    "peak_gbs": _measured_peaks().get("hbm_gbs"), "note": "HBM is 3-4 orders of magnitude off the critical path"},
            "launches_ms": [{"kernel": n, "R": r, "ms": round(ms, 4)} for n, r, ms in ktimes],
        }
        line = {
            "metric": "GCUPS", "value": cells_ref / (ms_step * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16x2 (int16 DPX lanes)", "data": "synthetic",
            "config": arm_config(n_cand, len(reads), codes.size, info["l2_bytes"], world),
            "sv_per_s": cand_total / (ms_step * 1e-3),
            "gcups_executed": cells_exec / (ms_step * 1e-3) / 1e9,
            "cells_reference_per_step": cells_ref, "cells_executed_per_step": cells_exec, "cells_pruned_per_step": cells_pruned,
            "reads_pruned_per_gpu_per_step": st["reads_pruned"],
            "supporting_reads_per_step": supp_total,
            "clocks": clocks,
            # both ways of driving the public call are measured every run (max over ranks each); `value` is the faster one:
            # with 8 ranks on one host the 16 concurrent uploads of the pipelined form can cost more than they hide
            "e2e": {"value": cells_ref / min(e2e_s, e2e_serial_s) / 1e9, "unit": "GCUPS", "ms_per_step": min(e2e_s, e2e_serial_s) * 1e3,
                    "sv_per_s": cand_total / min(e2e_s, e2e_serial_s), "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(st["d2h_bytes"]), "calls_in_flight": 2 if e2e_s <= e2e_serial_s else 1,
                    "pipelined": {"value": cells_ref / e2e_s / 1e9, "ms_per_step": e2e_s * 1e3,
                                  "note": "two nmsv_validate_batch calls in flight (two contexts, two host threads)"},
                    "serial": {"value": cells_ref / e2e_serial_s / 1e9, "ms_per_step": e2e_serial_s * 1e3,
                               "note": "one nmsv_validate_batch call at a time"}},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "roofline": roofline,
            "device": info["name"],
        }
    plan.close()
    if rank == 0 and world == 1:
        line["e2e_stage"] = stage_seam_e2e(eng, head_wl, head_order, n_cand, checked, supporting, len(records))
        log(f"[stage seam] {line['e2e_stage']['sv_per_s']:.0f} candidates/s through count_sread_from_seam "
            f"({line['e2e_stage']['seam_gb_per_s']:.2f} GB/s of seam TSV) vs {line['sv_per_s']:.0f}/s device resident")

    # ---- CPU leg on rank 0 at N = 1: the timed baseline, whose results are the oracle of the sampled parity check
    cpu_gcups = 80.0
    parity = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        job = OracleJob(pack, svs, reads, 2e10)
        dt, used, _ = job.run(threads)
        job = OracleJob(pack, svs, reads, max(2e10, job.cells / dt * 15.0))
        dt, used, res = job.run(threads)
        cpu_gcups = job.cells / dt / 1e9
        line["cpu_baseline"] = {
            "value": cpu_gcups, "unit": "GCUPS", "cores": used, "kind": "port",
            "sample": f"{job.n_svs} SV candidates of the same batch (tumor reads), {len(job.t_idx)} alignments with "
                      f"reversed begin pass, {job.cells / 1e9:.1f} G forward cells in {dt:.1f}s; engine oracle/nmsv_striped.c "
                      f"(AVX2 striped, 8-bit then 16-bit: parasail-equivalent port, parasail not installable)"}
        parity[WORKLOAD] = parity_validate(eng, pack, svs, reads, job, res, supporting)
        log(f"[config 2] parity {parity[WORKLOAD]['equal']}/{parity[WORKLOAD]['checked']} alignments equal the oracle")

    # ---- the other BASELINE configs
    configs = []
    if not args.no_configs:
        if world == 1 and rank == 0:
            configs.append({"baseline_config": 2, "workload": WORKLOAD, "sv_candidates": n_cand, "ms": ms_step,
                            "sv_per_s": line["sv_per_s"], "gcups": line["value"], "gcups_executed": line["gcups_executed"],
                            "roofline": {"frac": line["roofline"]["frac"]}, "note": "the headline measurement of this line"})
            for number, preset, n_sv, over in SHAPES:
                if args.no_cpu_baseline:
                    continue
                configs.append(shape_entry(torch, stream, eng, dpx_peak, number, preset, n_sv, over, cpu_gcups, 3.0, threads))
                parity[preset] = configs[-1]["parity_sampled"]
        sweep = pair_sweep_entry(torch, dist, dev, eng, rank, world, args.pairs, cpu_gcups, 0.0 if args.no_cpu_baseline else 3.0, threads)
        if rank == 0:
            configs.append(sweep)
            if "parity_sampled" in sweep:
                parity["pair_sweep"] = sweep["parity_sampled"]
        if world > 1:
            strong = _strong(torch, dist, dev, eng, rank, world, args)
            if rank == 0:
                line["strong"] = strong
    # ---- SURVEY 8(f) "next" row 2, driver-visible: sw_jump (breakpoint refinement) through nmsv_sw_jump_batch
    if rank == 0 and world == 1 and not args.no_configs:
        line["next_rows"] = {"sw_jump": sw_jump_entry(eng, check=0 if args.no_cpu_baseline else 40)}
        log(f"[sw_jump] {line['next_rows']['sw_jump']['abi_gcells_per_s']:.0f} Gcells/s through the C ABI, "
            f"parity {line['next_rows']['sw_jump']['parity_checked']} problems")
    if rank == 0:
        configs.sort(key=lambda c: c["baseline_config"])
        line["configs"] = configs
        if parity:
            line["parity_sampled"] = {"checked": sum(p["checked"] for p in parity.values()),
                                      "equal": sum(p["equal"] for p in parity.values()),
                                      "oracle": "oracle/nmsv_striped.c + the strand rule of ssw_check; the CUDA results of the same "
                                                "reads (a plan that computes every alignment) are compared field by field",
                                      "per_config": parity}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def sw_jump_entry(eng, n: int = 3000, check: int = 40) -> dict:
    """The 3000 locate_bp-shaped problems of bench_jump.py (reference smith_waterman.py:5-127) through the C ABI with host
    buffers; the first `check` problems are held to the C restatement of the reference's function (oracle/nmsv_jump_oracle.c,
    itself pinned on vectors made by the reference)."""
    import bench_jump
    from nanomonsv_b200 import smith_waterman as SW
    probs = bench_jump.problems(n)
    cells = sum((len(c) + 1) * (len(a) + len(b) + 2) for c, a, b in probs)
    SW.sw_jump_batch(probs, 1, 3, 3, 2, 10, engine=eng)          # grows the stream-ordered pool and the staging buffers once
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        got = SW.sw_jump_batch(probs, 1, 3, 3, 2, 10, engine=eng)
        dt = time.perf_counter() - t0
        best = SW.sw_jump_batch.last_abi_seconds if best is None else min(best, SW.sw_jump_batch.last_abi_seconds)
    if check:
        from oracle import oracle as O
        for (c, a, b), g in zip(probs[:check], got[:check]):
            if g != O.sw_jump(c, a, b, 1, 3, 3, 2, 10):
                raise SystemExit("sw_jump: the CUDA path differs from the oracle")
    return {"problems": n, "cells": cells, "abi_ms": best * 1e3, "abi_gcells_per_s": cells / best / 1e9,
            "python_mirror_problems_per_s": n / dt, "found": sum(g is not None for g in got), "parity_checked": check,
            "bound": "instruction issue of 8 resident warps per SM (13.6 / 18.4 instructions per H1 / H2 cell) and the length of "
                     "the longest problem; path planes 2 bits per cell: HBM below 5 % of its peak",
            "what": "two linear-gap DP matrices with a logarithmic jump + traceback per problem (contig x region1, contig x region2), "
                    "regions 60-900 bases, contigs up to 2.6 kb; host strings in, result tuples and traceback operations out"}


def _strong(torch, dist, dev, eng, rank, world, args):
    from bench_strong import strong_scaling
    return strong_scaling(torch, dist, dev, eng, rank, world, args.strong_svs)


def _traffic(cells_executed: float):
    """DRAM bytes of the forward launch. Not measurable inside a run (ncu replays kernels): the figure comes from the
    committed ncu capture of the same kernel on the 192-candidate batch, scaled by the executed cells (the traffic is
    the read bases, proportional to the cells at a fixed template length)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return t["dram_bytes_per_executed_cell"] * cells_executed, "from capture: " + t["source"] + ", scaled by executed cells"
    except Exception:   # noqa: BLE001
        return None, None


def _measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        return {"hbm_gbs": 6650.0, "source": "fallback"}


if __name__ == "__main__":
    main()
