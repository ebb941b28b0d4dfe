"""Multi-GPU sharding of a validation job: SV candidates are independent, so they are partitioned across the
ranks (one process per GPU) with no communication during compute; at the end the per-SV counts are merged by ONE
all-reduce and the supporting-read records by a count all-gather plus one padded all-gather (NCCL over NVLink on the
GPU box, gloo in the CPU tests).

The reference's own parallelism is the same shape: `--processes N` deals clusters round-robin into N shard files,
runs them in a process pool and concatenates the text outputs (run.py:324-379, :391-407).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def lpt_partition(costs: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first: heaviest unit to the currently lightest part. Deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * n_parts
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for i in order:
        p = min(range(n_parts), key=lambda q: (loads[q], q))
        parts[p].append(i)
        loads[p] += float(costs[i])
    for p in parts:
        p.sort()
    return parts


def round_robin_partition(n_units: int, n_parts: int) -> List[List[int]]:
    """The reference's dealing order (run.py:331-342)."""
    return [list(range(p, n_units, n_parts)) for p in range(n_parts)]


def sv_costs(svs: np.ndarray, reads: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Forward DP cells of every SV of a packed batch (both strands, identical templates counted once)."""
    from . import _lib
    read_len = lens[reads["seq"]].astype(np.int64)
    per_sv_bases = np.zeros(len(svs), dtype=np.int64)
    np.add.at(per_sv_bases, reads["sv"], read_len)
    rows = np.zeros(len(svs), dtype=np.int64)
    for i, sv in enumerate(svs):
        seen = set()
        for s in range(4):
            t = int(sv["tmpl"][s])
            if t == _lib.NMSV_NONE or t in seen:
                continue
            if (sv["flags"] & _lib.SV_SBND) and s in (1, 3):
                continue
            seen.add(t)
            rows[i] += int(lens[t])
    return 2 * rows * per_sv_bases


def merge_counts(n_total: int, local_indices: Sequence[int], local_checked: np.ndarray, local_supporting: np.ndarray,
                 group=None, device=None) -> np.ndarray:
    """Dense int32[n_total, 2] table (checked, supporting) merged over all ranks with one all-reduce(sum): every
    rank fills only the rows of its own SVs. Works on CPU tensors (gloo) and CUDA tensors (NCCL)."""
    import torch
    import torch.distributed as dist
    table = torch.zeros((n_total, 2), dtype=torch.int32, device=device or "cpu")
    if len(local_indices):
        idx = torch.as_tensor(np.asarray(local_indices, dtype=np.int64), device=table.device)
        vals = torch.as_tensor(np.stack([np.asarray(local_checked, dtype=np.int32),
                                         np.asarray(local_supporting, dtype=np.int32)], axis=1), device=table.device)
        table[idx] = vals
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table.cpu().numpy()


RECORD_COLS = 2 + 24     # SV index in the whole job, read ordinal inside the SV, 4 x (score, qstart, qend, tstart, tend, strand)


def records_to_rows(records: np.ndarray, svs: np.ndarray, global_sv_ids: Sequence[int]) -> np.ndarray:
    """nmsv_record[] of one rank's batch -> int32 rows that mean the same on every rank: the SV's index in the whole job
    and the read's ordinal inside its SV instead of batch-local indices."""
    rows = np.zeros((len(records), RECORD_COLS), dtype=np.int32)
    if len(records):
        sv_local = records["sv"].astype(np.int64)
        rows[:, 0] = np.asarray(global_sv_ids, dtype=np.int64)[sv_local]
        rows[:, 1] = records["read"].astype(np.int64) - svs["read_begin"][sv_local].astype(np.int64)
        rows[:, 2:] = records["aln"].view(np.int32).reshape(len(records), 24)
    return rows


def gather_record_rows(rows: np.ndarray, group=None, device=None) -> np.ndarray:
    """Supporting-read records of all ranks: one all-gather of the per-rank counts, then one padded all-gather of the
    int32 rows (on the device over NCCL, on CPU tensors over gloo). Returned sorted by (SV, read): identical on every
    rank and independent of the partition."""
    import torch
    import torch.distributed as dist
    rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, RECORD_COLS)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        dev = device or "cpu"
        n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, n, group=group)
        cap = max(int(counts.max().item()), 1)
        mine = torch.zeros((cap, RECORD_COLS), dtype=torch.int32, device=dev)
        if rows.shape[0]:
            mine[:rows.shape[0]] = torch.from_numpy(rows).to(dev)
        everything = torch.zeros((world, cap, RECORD_COLS), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(everything.view(-1), mine.view(-1), group=group)
        everything, counts = everything.cpu().numpy(), counts.cpu().numpy()
        rows = np.concatenate([everything[r, :int(counts[r])] for r in range(world)], axis=0)
    order = np.lexsort((rows[:, 1], rows[:, 0]))
    return rows[order]


def validate_sharded(costs: Sequence[float], run_local, rank: int, world: int, group=None, device=None,
                     partition: str = "lpt"):
    """Partition n = len(costs) SV candidates over `world` ranks, run `run_local(indices)` on this rank's share
    (-> (checked[int], supporting[int], record rows[int32, RECORD_COLS]) for those indices, column 0 = index in the
    whole job) and merge: returns the dense (n, 2) count table, the record rows of the whole job sorted by (SV, read)
    -- both identical on every rank -- and the partition. `run_local` is the GPU path in production (this rank's
    engine); the two collectives are the only communication."""
    n = len(costs)
    parts = lpt_partition(costs, world) if partition == "lpt" else round_robin_partition(n, world)
    mine = parts[rank]
    if mine:
        checked, supporting, rows = run_local(mine)
    else:
        checked, supporting, rows = [], [], np.zeros((0, RECORD_COLS), dtype=np.int32)
    table = merge_counts(n, mine, np.asarray(checked, dtype=np.int32), np.asarray(supporting, dtype=np.int32),
                         group=group, device=device)
    return table, gather_record_rows(rows, group=group, device=device), parts
