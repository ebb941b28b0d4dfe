"""Multi-GPU sharding of a validation job: SV candidates are independent, so they are partitioned across the
ranks (one process per GPU) with no communication during compute, and the per-SV counts are merged by ONE
collective at the end (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference's own parallelism is the same shape: `--processes N` deals clusters round-robin into N shard files,
runs them in a process pool and concatenates the text outputs (run.py:324-379, :391-407).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

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


def gather_records(local_lines: List[str], group=None) -> List[str]:
    """Variable-size supporting-read info lines of all ranks, in rank order (all_gather_object: host side)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local_lines)
    out: List[Optional[List[str]]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local_lines), group=group)
    return [ln for part in out for ln in (part or [])]


def validate_sharded(costs: Sequence[float], run_local, rank: int, world: int, group=None, device=None,
                     partition: str = "lpt"):
    """Partition n = len(costs) SV candidates over `world` ranks, run `run_local(indices)` on this rank's share
    (-> (checked[int], supporting[int], info_lines[str]) for those indices, in order) and merge: returns the dense
    (n, 2) count table and all info lines, identical on every rank. `run_local` is the GPU path in production
    (Alignment_counter on this rank's engine); the collective is the only communication."""
    n = len(costs)
    parts = lpt_partition(costs, world) if partition == "lpt" else round_robin_partition(n, world)
    mine = parts[rank]
    checked, supporting, lines = run_local(mine) if mine else ([], [], [])
    table = merge_counts(n, mine, np.asarray(checked, dtype=np.int32), np.asarray(supporting, dtype=np.int32),
                         group=group, device=device)
    return table, gather_records(list(lines), group=group), parts
