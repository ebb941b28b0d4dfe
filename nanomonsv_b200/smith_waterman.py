"""GPU-backed drop-in for nanomonsv/smith_waterman.py (reference v0.8.0): `sw_jump` with the reference's signature
and return value (smith_waterman.py:5-127), called by locate_bp.get_refined_bp (locate_bp.py:37,67), plus a batched
form. The DP and the traceback run in CUDA (csrc/nmsv_jump.cuh) behind nmsv_sw_jump_batch; only the two alignment
strings of the return value are assembled here from the traceback operations."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .engine import Engine, default_engine


_DASH = np.uint8(ord("-"))


def _walk(seq_c: np.ndarray, seq_r: np.ndarray, i_end: int, j_end: int, ops: np.ndarray) -> Tuple[bytes, bytes]:
    """Alignment rows of one traceback (reference smith_waterman.py:71-125), vectorised: operation k consumes a contig
    base when its code is 1, 2 (or 4, the jump cell) and a region base when it is 1, 3 (or 4)."""
    if len(ops) == 0:
        return b"", b""
    uses_c = (ops == 1) | (ops == 2) | (ops == 4)
    uses_r = (ops == 1) | (ops == 3) | (ops == 4)
    moves_c = (ops == 1) | (ops == 2)
    moves_r = (ops == 1) | (ops == 3)
    i = i_end - np.concatenate(([0], np.cumsum(moves_c[:-1])))
    j = j_end - np.concatenate(([0], np.cumsum(moves_r[:-1])))
    c = np.where(uses_c, seq_c[np.clip(i - 1, 0, len(seq_c) - 1)], _DASH)
    g = np.where(uses_r, seq_r[np.clip(j - 1, 0, len(seq_r) - 1)], _DASH)
    return c[::-1].tobytes(), g[::-1].tobytes()


def _strings(contig: str, region1: str, region2: str, r, ops: np.ndarray) -> Tuple[str, str]:
    cb = np.frombuffer(contig.encode("latin-1"), dtype=np.uint8)
    n2, n1 = int(r["n_ops2"]), int(r["n_ops1"])
    c2, g2 = _walk(cb, np.frombuffer(region2.encode("latin-1"), dtype=np.uint8), int(r["i2_end"]), int(r["j2_end"]), ops[:n2])
    c1, g1 = _walk(cb, np.frombuffer(region1.encode("latin-1"), dtype=np.uint8), int(r["i1_end"]), int(r["j1_end"]), ops[n2:n2 + n1])
    return (c1 + b" " + c2).decode("latin-1"), (g1 + b" " + g2).decode("latin-1")


def sw_jump_batch(problems: Sequence[Tuple[str, str, str]], match_score: int = 1, mismatch_penalty: int = 2,
                  gap_cost: int = 2, jump_cost: int = 2, minimum_score: int = 10,
                  engine: Optional[Engine] = None) -> List[Optional[tuple]]:
    """[(contig, region1_seq, region2_seq), ...] -> list of sw_jump return values (None or the 6-tuple)."""
    if not problems:
        return []
    eng = engine or default_engine()
    chunks, lens = [], []
    for trio in problems:
        for sq in trio:
            if not sq:
                raise ValueError("sw_jump: empty sequence")
            chunks.append(np.frombuffer(sq.encode("latin-1"), dtype=np.uint8))
            lens.append(len(sq))
    lens = np.asarray(lens, dtype=np.uint32)
    offsets = np.zeros(len(lens), dtype=np.uint64)
    np.cumsum(lens[:-1].astype(np.uint64), out=offsets[1:])
    data = np.ascontiguousarray(np.concatenate(chunks))
    n = len(problems)
    idx = np.arange(3 * n, dtype=np.uint32)
    cidx, aidx, bidx = (np.ascontiguousarray(idx[k::3]) for k in range(3))
    per_ops = 2 * lens[0::3].astype(np.int64) + lens[1::3] + lens[2::3] + 4
    ops_off = np.concatenate(([0], np.cumsum(per_ops)))
    # the operations array is kept on the engine from call to call (a fresh array is fresh pages: the library's copy into
    # it then takes a page fault per 4 KB -- 2 ms of a 12 ms call for 3000 problems); only the first n_ops2 + n_ops1 bytes
    # of a problem's segment are meaningful
    ops = getattr(eng, "_jump_ops", None)
    if ops is None or ops.size < int(ops_off[-1]):
        ops = np.zeros(max(int(ops_off[-1]) * 5 // 4, 1 << 20), dtype=np.uint8)
        eng._jump_ops = ops
    out = np.zeros(n, dtype=_lib.JUMP_RESULT_DTYPE)
    with np.errstate(divide="ignore"):
        logtab = np.log(np.arange(int(lens[0::3].max()) + 1))      # numpy's log, like the reference (:40)
    prm = _lib.JumpParams(match_score, mismatch_penalty, gap_cost, jump_cost, minimum_score)
    import time
    t0 = time.perf_counter()
    eng._check(eng._L.nmsv_sw_jump_batch(eng._h, data.ctypes.data, data.size, offsets.ctypes.data, lens.ctypes.data,
                                         len(lens), cidx.ctypes.data, aidx.ctypes.data, bidx.ctypes.data, n,
                                         ctypes.byref(prm), logtab.ctypes.data, len(logtab), out.ctypes.data,
                                         ops.ctypes.data, ops.size))
    sw_jump_batch.last_abi_seconds = time.perf_counter() - t0      # the C-ABI call alone (H2D, kernels, D2H), for bench_jump.py
    res: List[Optional[tuple]] = []
    for k, (contig, r1, r2) in enumerate(problems):
        r = out[k]
        if not r["found"]:
            res.append(None)
            continue
        cm, rm = _strings(contig, r1, r2, r, ops[int(ops_off[k]):int(ops_off[k + 1])])
        res.append((int(r["score"]), (int(r["i1_start"]), int(r["i1_end"]), int(r["i2_start"]), int(r["i2_end"])),
                    (int(r["j1_start"]), int(r["j1_end"])), (int(r["j2_start"]), int(r["j2_end"])), cm, rm))
    return res


def sw_jump(contig, region1_seq, region2_seq, match_score=1, mismatch_penalty=2, gap_cost=2, jump_cost=2,
            minimum_score=10, engine: Optional[Engine] = None):
    """Same signature and return value as the reference's sw_jump (a batch of one; use sw_jump_batch for throughput)."""
    return sw_jump_batch([(contig, region1_seq, region2_seq)], match_score, mismatch_penalty, gap_cost, jump_cost,
                         minimum_score, engine)[0]
