"""GPU-backed drop-in for nanomonsv/smith_waterman.py (reference v0.8.0): `sw_jump` with the reference's signature
and return value (smith_waterman.py:5-127), called by locate_bp.get_refined_bp (locate_bp.py:37,67), plus a batched
form. The DP and the traceback run in CUDA (csrc/nmsv_jump.cuh) behind nmsv_sw_jump_batch; the two alignment
strings of the return value are written out from the traceback operations by the library's host-side companion
nmsv_sw_jump_strings."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .engine import Engine, default_engine


def sw_jump_batch(problems: Sequence[Tuple[str, str, str]], match_score: int = 1, mismatch_penalty: int = 2,
                  gap_cost: int = 2, jump_cost: int = 2, minimum_score: int = 10,
                  engine: Optional[Engine] = None) -> List[Optional[tuple]]:
    """[(contig, region1_seq, region2_seq), ...] -> list of sw_jump return values (None or the 6-tuple)."""
    if not problems:
        return []
    eng = engine or default_engine()
    enc = [sq.encode("latin-1") for trio in problems for sq in trio]
    lens = np.fromiter(map(len, enc), dtype=np.uint32, count=len(enc))
    if not lens.all():
        raise ValueError("sw_jump: empty sequence")
    offsets = np.zeros(len(lens), dtype=np.uint64)
    np.cumsum(lens[:-1].astype(np.uint64), out=offsets[1:])
    data = np.frombuffer(b"".join(enc), dtype=np.uint8)
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
    # the two alignment rows per problem (reference :61-125), written by the library at the offsets of the operations
    rows = getattr(eng, "_jump_rows", None)
    if rows is None or rows[0].size < ops.size:
        rows = (np.zeros(ops.size, dtype=np.uint8), np.zeros(ops.size, dtype=np.uint8))
        eng._jump_rows = rows
    row_off = np.ascontiguousarray(ops_off[:-1].astype(np.uint64))
    eng._check(eng._L.nmsv_sw_jump_strings(data.ctypes.data, offsets.ctypes.data, lens.ctypes.data, cidx.ctypes.data,
                                           aidx.ctypes.data, bidx.ctypes.data, n, out.ctypes.data, ops.ctypes.data,
                                           row_off.ctypes.data, rows[0].ctypes.data, rows[1].ctypes.data))
    cols = {f: out[f].tolist() for f in out.dtype.names}
    starts = row_off.tolist()
    crow, rrow = rows[0].data, rows[1].data
    res: List[Optional[tuple]] = []
    for k in range(n):
        if not cols["found"][k]:
            res.append(None)
            continue
        a = starts[k]
        b = a + cols["n_ops1"][k] + 1 + cols["n_ops2"][k]
        res.append((cols["score"][k], (cols["i1_start"][k], cols["i1_end"][k], cols["i2_start"][k], cols["i2_end"][k]),
                    (cols["j1_start"][k], cols["j1_end"][k]), (cols["j2_start"][k], cols["j2_end"][k]),
                    str(crow[a:b], "latin-1"), str(rrow[a:b], "latin-1")))
    return res


def sw_jump(contig, region1_seq, region2_seq, match_score=1, mismatch_penalty=2, gap_cost=2, jump_cost=2,
            minimum_score=10, engine: Optional[Engine] = None):
    """Same signature and return value as the reference's sw_jump (a batch of one; use sw_jump_batch for throughput)."""
    return sw_jump_batch([(contig, region1_seq, region2_seq)], match_score, mismatch_penalty, gap_cost, jump_cost,
                         minimum_score, engine)[0]
