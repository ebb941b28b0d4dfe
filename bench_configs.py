#!/usr/bin/env python
"""Secondary measurement: the wavefront kernel on the SHAPES of the other BASELINE.json configs (bench.py measures the
headline configuration, configs[1]). One GPU, inputs resident in HBM, CUDA events inside the library (per-launch
times) and around the whole run; every configuration is scaled down in the number of candidates only -- template
lengths, read-length distributions and the mix of SV types are the presets of nanomonsv_b200/synth.py.

    python bench_configs.py [--out profiles/r01_configs.json]

  testdata_shaped   config 1 stand-in: m = 400 (strip class R = 13, one tile), reads 5-39 kb
  colo829_shaped    config 2 (the bench.py workload) at a small size, for comparison: m = 2000, R = 16, 4 tiles
  insertion_heavy   config 3: m = 6000 (12 tiles), var1 != var2
  single_breakend   config 4: 601- and 400-base templates (R = 19 and R = 13 in one plan), long soft-clipped tails
  pair_sweep        config 5: pair API, m = 400, read length uniform in 1-50 kb, half of the reads carry the template
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--only", default=None, help="comma-separated subset of the configurations")
    args = ap.parse_args()

    import torch
    from nanomonsv_b200 import Engine, SeqPack, synth

    eng = Engine(0)
    results = []

    def timed(fn):
        best = None
        for _ in range(args.repeat + 1):      # first run = warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, out

    plans = [("testdata_shaped", dict(n_sv=60)), ("colo829_shaped", dict(n_sv=40, contig_len=2_000_000, n_contigs=2)),
             ("insertion_heavy", dict(n_sv=int(os.environ.get("NMSV_BC_INS", "24")), contig_len=2_000_000, n_contigs=2)),
             ("single_breakend", dict(n_sv=300, contig_len=2_000_000, n_contigs=2))]
    only = set(args.only.split(",")) if args.only else None
    for name, kw in plans:
        if only and name not in only:
            continue
        cfg = dataclasses.replace(synth.PRESETS[name], **kw)
        t0 = time.time()
        wl = synth.make_workload(cfg)
        entry = {"config": name, "candidates": cfg.n_sv, "validate_sequence_length": cfg.validate_sequence_length}
        cells_ref = cells_exec = 0
        ms = 0.0
        n_reads = supp = 0
        launches = []
        for sample in ("tumor", "control"):
            pack, svs, reads = wl.to_batch(sample, var_read_min_mapq=0)
            plan = eng.plan(pack, svs, reads)
            plan.set_profiling(True)
            dt, _ = timed(plan.run)
            _, supporting, _ = plan.download()
            st = plan.stats()
            cells_ref += st["cells_reference"]; cells_exec += st["cells_executed"]
            ms += dt * 1e3
            n_reads += len(reads); supp += int(supporting.sum())
            launches += [(k, r, round(m, 3)) for k, r, m in plan.kernel_times() if m > 0.05]
            plan.close()
        entry.update(reads=n_reads, supporting_reads=supp, ms=round(ms, 2),
                     gcups_reference_equivalent=round(cells_ref / ms / 1e6, 1), gcups_executed=round(cells_exec / ms / 1e6, 1),
                     sv_per_s=round(cfg.n_sv / (ms * 1e-3), 1), launches_ms=launches, build_s=round(time.time() - t0, 1))
        print(json.dumps(entry), flush=True)
        results.append(entry)

    if only and "pair_sweep" not in only:
        if args.out:
            json.dump({"device": eng.device_info(), "results": results}, open(args.out, "w"), indent=1)
        return
    # config 5: the pair API (ssw_check semantics: both strands + begin pass)
    seqs, t_idx, r_idx = synth.make_pair_sweep(20000, m=400, n_min=1000, n_max=50000, pool_reads=4000, seed=5)
    pack = SeqPack()
    for s in seqs:
        pack.add_codes(s)
    cells = float(sum(2 * 400 * len(seqs[r]) for r in r_idx))
    dt, chk = timed(lambda: eng.ssw_check_batch(pack, t_idx, r_idx))
    entry = {"config": "pair_sweep", "pairs": len(t_idx), "m": 400, "ms": round(dt * 1e3, 2),
             "gcups_reference_equivalent": round(cells / dt / 1e9, 1),
             "hits": int((chk["score"] > 480).sum()),
             "note": "through nmsv_ssw_check_batch with host buffers: includes H2D of the sequences and D2H of the tuples"}
    print(json.dumps(entry), flush=True)
    results.append(entry)

    if args.out:
        json.dump({"device": eng.device_info(), "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
