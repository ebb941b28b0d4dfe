"""Strong scaling of the product's own multi-GPU path (bench.py --gpus N, N > 1, `strong` section of the JSON line).

ONE fixed validation job -- the same SV candidates on every rank (same seed), `repeat` instances of each so that the job
is large enough for 8 GPUs without minutes of synthetic-data generation -- is partitioned over the ranks with
nanomonsv_b200.shard (forward-cell costs, longest-processing-time-first), every rank validates ITS share on its GPU
(Engine plan, device resident), the per-SV counts are merged by one NCCL all-reduce and the supporting-read records by a
count all-gather + one padded all-gather of int32 rows on the device. Rank 0 then runs the whole job alone (N = 1) and
compares: count table and record rows must be identical. Times are CUDA-event device times of the per-rank work (max
over ranks) plus the wall time of the collectives.

The reference's own sharding is the same shape: clusters dealt round-robin to `--processes` workers, text outputs
concatenated (run.py:324-379, :391-407).
"""
from __future__ import annotations

import time

import numpy as np


def _sub_batch(pack_arrays, svs, reads, picks):
    """The SV instances `picks` (indices into svs, repeats allowed) as a batch of their own over the same sequence pack."""
    from nanomonsv_b200 import _lib
    from nanomonsv_b200.engine import SeqPack
    codes, offsets, lens = pack_arrays
    out_sv = np.zeros(len(picks), dtype=_lib.SV_DTYPE)
    rows = []
    pos = 0
    for k, i in enumerate(picks):
        sv = svs[i].copy()
        rb, re_ = int(sv["read_begin"]), int(sv["read_end"])
        r = reads[rb:re_].copy()
        r["sv"] = k
        sv["read_begin"], sv["read_end"] = pos, pos + len(r)
        pos += len(r)
        out_sv[k] = sv
        rows.append(r)
    out_rd = np.concatenate(rows) if rows else np.zeros(0, dtype=_lib.READ_DTYPE)
    return SeqPack.from_arrays(codes, offsets, lens), out_sv, out_rd


def strong_scaling(torch, dist, dev, eng, rank, world, n_svs, repeat=4):
    import bench
    from nanomonsv_b200 import shard
    cfg, pack, svs, reads = bench.build_batch(n_svs, 0, None, budget=None)      # rank-independent seed: the SAME job everywhere
    arrays = pack.finalize()
    lens = arrays[2]
    base_costs = shard.sv_costs(svs, reads, lens).astype(np.float64)
    instances = np.tile(np.arange(len(svs)), repeat)                             # instance j is a copy of SV j % len(svs)
    costs = base_costs[instances]
    stream = torch.cuda.current_stream()

    def run_share(picks, steps=2):
        spack, ssvs, sreads = _sub_batch(arrays, svs, reads, [int(instances[j]) for j in picks])
        plan = eng.plan(spack, ssvs, sreads)
        plan.run()                                                               # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            plan.run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        checked, supporting, records = plan.download()
        st = plan.stats()
        plan.close()
        return ms, checked, supporting, shard.records_to_rows(records, ssvs, picks), st

    timing = {}

    def run_local(picks):
        ms, checked, supporting, rows, st = run_share(picks)
        timing.update(ms=ms, cells=st["cells_executed"], cells_ref=st["cells_reference"])
        return checked, supporting, rows

    dist.barrier()
    t0 = time.perf_counter()
    table, rows, parts = shard.validate_sharded(costs, run_local, rank, world, device=dev)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    red = torch.tensor([timing.get("ms", 0.0)], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(timing.get("cells", 0)), float(timing.get("cells_ref", 0))], dtype=torch.float64, device=dev)
    mn = red.clone()
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(tot)
    # the collectives on their own (counts all-reduce + record gather), timed once more on the merged data
    dist.barrier()
    t0 = time.perf_counter()
    shard.merge_counts(len(costs), parts[rank], table[parts[rank], 0], table[parts[rank], 1], device=dev)
    shard.gather_record_rows(rows[np.isin(rows[:, 0], parts[rank])], device=dev)
    torch.cuda.synchronize()
    coll_s = time.perf_counter() - t0
    out = None
    if rank == 0:
        ms1, c1, s1, rows1, st1 = run_share(list(range(len(costs))))
        order = np.lexsort((rows1[:, 1], rows1[:, 0]))
        same = bool((table[:, 0] == c1).all() and (table[:, 1] == s1).all() and (rows == rows1[order]).all())
        if not same:
            raise SystemExit("strong-scaling job: the merged N-rank results differ from the single-GPU run")
        loads = [float(costs[p].sum()) for p in parts]
        t_n = float(red[0])
        out = {"job": f"{len(svs) // 2} candidates x (tumor, control) x {repeat} instances = {len(costs)} (SV, sample) units of "
                      f"the headline workload, identical on every rank",
               "n_gpus": world, "partition": "shard.lpt_partition by forward cells",
               "ms_n_gpus": t_n, "ms_1_gpu": ms1, "speedup": ms1 / t_n, "efficiency": ms1 / t_n / world,
               "ms_fastest_rank": float(mn[0]), "lpt_imbalance": max(loads) / (sum(loads) / world) - 1.0,
               "gcups_executed": float(tot[0]) / t_n / 1e6, "gcups": float(tot[1]) / t_n / 1e6,
               "collectives_ms": coll_s * 1e3, "wall_ms_incl_planning_and_gather": wall * 1e3,
               "records": int(len(rows)), "identical_to_single_gpu": same,
               "limiter": "the end of each rank's launch (its longest reads run at low occupancy) does not shrink with N; "
                          "LPT imbalance and the two collectives are the figures beside it"}
        bench.log(f"[strong] {world} GPUs: {t_n:.1f} ms vs {ms1:.1f} ms on one -> efficiency {out['efficiency']:.3f}, "
                  f"imbalance {out['lpt_imbalance']:.4f}, collectives {coll_s * 1e3:.2f} ms")
    dist.barrier()
    return out
