#!/usr/bin/env python3
"""bench.py -- supporting-read validation throughput (GCUPS, SV candidates/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # host-CPU "parasail-equivalent" port

Workload = BASELINE.json configs[1] ("synthetic COLO829-shaped": canonical SV candidates, 30x tumor + 30x control
ONT-like reads ~10 kb at ~8 % error, +-1 kb templates -> m = 2000). One STEP validates one batch of
--svs-per-gpu candidates (tumor AND control reads) per GPU; the 20k-candidate job of the config is that step
repeated, so GCUPS and candidates/s do not depend on the batch count. Weak scaling: every rank gets its own
batch (own seed), no data-path collective; for N > 1 the per-SV counts are merged with one NCCL all-reduce per step.

value  = reference-equivalent forward DP cells (sum 2*m*n over every alignment parasail would run) / device time,
         inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e    = the same cells / wall time of nmsv_validate_batch called with pinned HOST buffers (H2D, planning,
         kernels, D2H inside the timed region), two calls in flight (two contexts, two host threads) so that the
         upload of one step overlaps the kernels of the previous one; `e2e.serial` is the same with one call at a time.
Identical templates of one SV (var1 == var2 when there is no insertion) are aligned once on the GPU; the cells
actually executed are reported as `cells_executed` and are what the roofline fraction is computed from.
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
# mean EXECUTED forward cells of one synthetic candidate (tumor + control reads, both strands, m = 2000, identical
# templates counted once), measured on the generator; the per-GPU batch is cut at --svs-per-gpu times this budget so
# that every rank does equal device work
CELLS_PER_CANDIDATE = 7.926e9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout. Native libraries (NCCL's version banner, ...) write to fd 1 directly, so
# fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def build_batch(n_sv: int, rank: int, args):
    """One step's batch for one GPU: tumor reads then control reads of the same n_sv candidates."""
    from nanomonsv_b200 import synth
    from nanomonsv_b200.batch import BatchBuilder
    from nanomonsv_b200.count_sread_by_alignment import canonical_templates
    cfg = dataclasses.replace(synth.PRESETS[WORKLOAD], seed=synth.PRESETS[WORKLOAD].seed + 7919 * rank,
                              contig_len=4_000_000, n_contigs=4)
    t0 = time.time()
    # weak scaling needs equal work per rank: draw some spare candidates and keep the first ones (in generation
    # order) whose reference-equivalent cell count reaches the per-GPU budget of n_sv average candidates
    spare = max(8, n_sv // 5)
    wl = synth.make_workload(cfg, n_sv + spare)
    L = cfg.validate_sequence_length
    per_sv = []
    for i, sv in enumerate(wl.svs):
        bases = sum(len(r.codes) for smp in ("tumor", "control") for r in wl.samples[smp][i])
        v1, v2, r1, r2, short = canonical_templates(sv.key, wl.fasta, L)
        distinct = {v1, v2} | ({r1, r2} if short else set())     # identical templates are aligned once on the GPU
        per_sv.append(2 * sum(len(t) for t in distinct) * bases)
    budget = CELLS_PER_CANDIDATE * n_sv
    keep, acc = [], 0
    for i, c in enumerate(per_sv):
        if acc >= budget:
            break
        keep.append(i); acc += c
    keep_set = set(keep)
    bb = BatchBuilder(score_ratio_thres=1.2, var_read_min_mapq=0)
    for sample in ("tumor", "control"):
        for i in [j for j in wl.sv_order() if j in keep_set]:
            var1, var2, ref1, ref2, short = canonical_templates(wl.svs[i].key, wl.fasta, L)
            bb.add_sv([var1, var2, ref1 if short else None, ref2 if short else None], False, short)
            for rd in wl.samples[sample][i]:
                bb.add_read(rd.codes, rd.mapq1, rd.mapq2)
    pack, svs, reads = bb.finalize()
    log(f"[rank {rank}] synthetic batch: {len(keep)} candidates x (tumor, control), {len(reads)} reads, "
        f"{pack.finalize()[0].size / 1e6:.0f} MB of bases, built in {time.time() - t0:.1f}s")
    return cfg, pack, svs, reads


def cpu_sample(pack, svs, reads, target_cells: float):
    """A bounded sample of the batch for the CPU leg: the first SVs (tumor part) until target_cells forward cells.
    Every alignment the reference would run is listed: each present template and its reverse complement against
    every read of the SV (no sharing of identical templates, like the reference)."""
    from nanomonsv_b200 import _lib
    from nanomonsv_b200.engine import SeqPack
    codes, offsets, lens = pack.finalize()
    comp = np.array([3, 2, 1, 0, 4], dtype=np.uint8)
    sp = SeqPack()
    t_idx, r_idx = [], []
    cells = 0
    n_used = 0
    for i, sv in enumerate(svs):
        if cells >= target_cells:
            break
        rd_idx = {}
        for r in range(int(sv["read_begin"]), int(sv["read_end"])):
            s = int(reads["seq"][r])
            rd_idx[r] = sp.add_codes(codes[int(offsets[s]):int(offsets[s]) + int(lens[s])])
        for slot in range(4):
            t = int(sv["tmpl"][slot])
            if t == _lib.NMSV_NONE:
                continue
            tc = codes[int(offsets[t]):int(offsets[t]) + int(lens[t])]
            f = sp.add_codes(tc)
            rc = sp.add_codes(comp[tc[::-1]])
            for r, ri in rd_idx.items():
                t_idx += [f, rc]
                r_idx += [ri, ri]
                cells += 2 * int(lens[t]) * int(lens[int(reads["seq"][r])])
        n_used += 1
    return sp, np.asarray(t_idx, dtype=np.uint32), np.asarray(r_idx, dtype=np.uint32), cells, n_used


def run_cpu(sp, t_idx, r_idx, threads: int):
    from oracle import oracle as O
    codes, offsets, lens = sp.finalize()
    t0 = time.perf_counter()
    _, used = O.ssw_batch(codes, offsets, lens, t_idx, r_idx, engine="striped", threads=threads)
    return time.perf_counter() - t0, used


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
def reference_arm(args, rank: int):
    if rank != 0:
        return
    cfg, pack, svs, reads = build_batch(min(args.svs_per_gpu, 16), 0, args)
    threads = os.cpu_count() or 1
    # calibrate on a small sample, then size one step to ~args.cpu_seconds
    sp, t, r, cells, _ = cpu_sample(pack, svs, reads, 2e10)
    dt, used = run_cpu(sp, t, r, threads)
    target = max(2e10, cells / dt * args.cpu_seconds)
    sp, t, r, cells, n_used = cpu_sample(pack, svs, reads, target)
    for _ in range(args.warmup):
        run_cpu(sp, t, r, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, used = run_cpu(sp, t, r, threads)
    dt = (time.perf_counter() - t0) / args.steps
    gcups = cells / dt / 1e9
    sample = (f"{n_used} of the batch's SV candidates (tumor reads), {len(t)} alignments incl. reverse-complement strand "
              f"and reversed begin pass, {cells / 1e9:.1f} G forward cells per step")
    line = {"impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8->int16 striped SIMD (AVX2)", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: m=2000 templates, ~10 kb reads, 8% error; bounded sample per step",
                       "engine": "oracle/nmsv_striped.c (parasail-equivalent port; parasail itself is not installable here)"},
            "sv_per_s": n_used / 2 / dt,
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--svs-per-gpu", type=int, default=int(os.environ.get("NMSV_BENCH_SVS", "192")))
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU work per step of the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    codes, offsets, lens = pack.finalize()
    n_cand = len(svs) // 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident hot path
    plan = eng.plan(pack, svs, reads)
    plan.set_profiling(True)
    st = plan.stats()
    # the path's only exchange: ONE merge of the per-GPU supporting-read count tables at the end of the job
    # (NCCL all-reduce of a dense int32 table over NVLink); the job here = the timed steps
    cap = len(svs) + 64
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
    ktimes = plan.kernel_times()                      # launches of the last timed step
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
    assert all((s_ == supporting).all() and len(r_) == len(records) for _, s_, r_ in results), "pipelined e2e results differ"
    clocks = sampler.stop()
    assert (s_e2e == supporting).all() and len(r_e2e) == len(records), "e2e and device-resident results differ"

    # ---- max over ranks, totals over ranks
    red = torch.tensor([ms_step, e2e_s, e2e_serial_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([st["cells_reference"], st["cells_executed"], n_cand, int(supporting.sum())],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_step, e2e_s, e2e_serial_s = float(red[0]), float(red[1]), float(red[2])
    cells_ref, cells_exec, cand_total, supp_total = [float(x) for x in tot]

    if rank == 0:
        fwd = [(r, ms) for name, r, ms in ktimes if name == "sw_forward"]
        dom_r, dom_ms = max(fwd, key=lambda x: x[1])
        step_kernel_ms = sum(ms for _, _, ms in ktimes)
        achieved = st["cells_executed"] * DPX_PER_CELL / (dom_ms * 1e-3)     # forward cells all run in this launch
        h2d = st["h2d_bytes"]
        roofline = {
            "bound": "int_alu (DPX issue rate; not hbm, not tensor)",
            "kernel": f"sw_wavefront_kernel<{dom_r}> (forward launch; it also drains the fused begin pass, whose cells are not counted)",
            "achieved": achieved / 1e12, "peak": dpx_peak / 1e12, "unit": "T DPX lane-inst/s",
            "frac": achieved / dpx_peak,
            "peak_source": "measured live: register-only VIADDMNMX.S16x2 loop on this GPU (nmsv_measure_dpx_peak)",
            "dpx_lane_inst_per_cell": DPX_PER_CELL,
            "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / step_kernel_ms,
            "gcups_kernel": st["cells_executed"] / (dom_ms * 1e-3) / 1e9,
            "traffic": _traffic(st["cells_executed"])[0],            # DRAM bytes per launch (ncu), or null
            "traffic_source": _traffic(st["cells_executed"])[1],
            "hbm": {"algorithmic_bytes": int(codes.size), "achieved_gbs": codes.size / (dom_ms * 1e-3) / 1e9,
                    "peak_gbs": _measured_peaks().get("hbm_gbs"), "note": "HBM is 3-4 orders of magnitude off the critical path"},
            "launches_ms": [{"kernel": n, "R": r, "ms": round(ms, 4)} for n, r, ms in ktimes],
        }
        line = {
            "metric": "GCUPS", "value": cells_ref / (ms_step * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16x2 (int16 DPX lanes)", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} (BASELINE.json configs[1]): canonical SV candidates, 30x tumor + 30x control "
                                   f"ONT-like reads ~10 kb, 8% error, m=2000 templates",
                       "sv_candidates_per_gpu_per_step": n_cand, "reads_per_gpu_per_step": int(len(reads)),
                       "read_bases_per_gpu_mb": round(codes.size / 1e6, 1),
                       "l2": f"inputs ({codes.size / 1e6:.0f} MB of read bases per GPU) exceed the {info['l2_bytes'] / 1e6:.0f} MB L2",
                       "parallelism": f"sv-shards x{world}, equal cell budget per GPU, one NCCL all-reduce of the count tables at the end of the job" if world > 1 else "single GPU"},
            "sv_per_s": cand_total / (ms_step * 1e-3),
            "gcups_executed": cells_exec / (ms_step * 1e-3) / 1e9,
            "cells_reference_per_step": cells_ref, "cells_executed_per_step": cells_exec,
            "supporting_reads_per_step": supp_total,
            "clocks": clocks,
            "e2e": {"value": cells_ref / e2e_s / 1e9, "unit": "GCUPS", "ms_per_step": e2e_s * 1e3,
                    "sv_per_s": cand_total / e2e_s, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(st["d2h_bytes"]), "calls_in_flight": 2,
                    "serial": {"value": cells_ref / e2e_serial_s / 1e9, "ms_per_step": e2e_serial_s * 1e3,
                               "note": "one nmsv_validate_batch call at a time"}},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "roofline": roofline,
            "device": info["name"],
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sp, t, r, cells, _ = cpu_sample(pack, svs, reads, 2e10)
            dt, used = run_cpu(sp, t, r, threads)
            sp, t, r, cells, n_used = cpu_sample(pack, svs, reads, max(2e10, cells / dt * 15.0))
            dt, used = run_cpu(sp, t, r, threads)
            line["cpu_baseline"] = {
                "value": cells / dt / 1e9, "unit": "GCUPS", "cores": used, "kind": "port",
                "sample": f"{n_used} SV candidates of the same batch (tumor reads), {len(t)} alignments with "
                          f"reversed begin pass, {cells / 1e9:.1f} G forward cells in {dt:.1f}s; engine oracle/nmsv_striped.c "
                          f"(AVX2 striped, 8-bit then 16-bit: parasail-equivalent port, parasail not installable)"}
        emit(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def _traffic(cells_executed: float):
    """DRAM bytes of the forward launch: ncu capture under profiles/, scaled by the executed cells of this batch
    (the traffic is the per-column tile-boundary spill, proportional to the cells). None if the capture is absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_final_traffic.json")))
        return t["dram_bytes_per_executed_cell"] * cells_executed, t["source"] + ", scaled by executed cells"
    except Exception:   # noqa: BLE001
        return None, None


def _measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        return {"hbm_gbs": 6650.0, "source": "fallback"}


if __name__ == "__main__":
    main()
