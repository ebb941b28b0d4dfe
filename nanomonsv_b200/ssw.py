"""Operator seam: a parasail-shaped facade over the CUDA engine.

The reference calls `parasail.matrix_create("ACGT", 2, -2)` and `parasail.ssw(s1, s2, open, extend, matrix)`
(count_sread_by_alignment.py:139,148-149; also post_proc.py:118-120, locate_bp_sbnd.py:28-29,
generate_consensus.py:119) and reads `.score1 .read_begin1 .read_end1 .ref_begin1 .ref_end1` from the result.
`ssw` / `ssw_batch` here return the same fields, computed on the GPU; a single call is a batch of one (use
`ssw_batch` for throughput).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from .engine import Engine, SeqPack, default_engine


@dataclass(frozen=True)
class Matrix:
    alphabet: str
    match: int
    mismatch: int


def matrix_create(alphabet: str, match: int, mismatch: int) -> Matrix:
    """Only the nucleotide alphabet the reference uses is supported by the byte encoding (ACGT + wildcard)."""
    if alphabet.upper() != "ACGT":
        raise ValueError("the GPU engine encodes ACGT (case-insensitive) plus a zero-scoring wildcard only")
    return Matrix(alphabet, int(match), int(mismatch))


@dataclass(frozen=True)
class SswResult:
    score1: int
    ref_begin1: int
    ref_end1: int
    read_begin1: int
    read_end1: int


def ssw_batch(pairs: Sequence[Tuple[str, str]], open: int, extend: int, matrix: Matrix,
              engine: Optional[Engine] = None) -> List[Optional[SswResult]]:
    """[(s1, s2), ...] -> results; None where parasail returns NULL: the SCORE is not representable in 16 bits (the
    engine reports that per pair as score1 = -1; the lengths of s1 and s2 do not matter, generate_consensus.py:119
    passes whole reads as s1)."""
    eng = engine or default_engine()
    out: List[Optional[SswResult]] = [None] * len(pairs)
    if not pairs:
        return out
    pack = SeqPack(dedup=True)
    t_idx, r_idx = [], []
    for s1, s2 in pairs:
        if not s1 or not s2:
            raise ValueError("empty sequence")
        t_idx.append(pack.add(s1)); r_idx.append(pack.add(s2))
    res = eng.ssw_batch(pack, t_idx, r_idx, (matrix.match, matrix.mismatch, open, extend), want_begin=True)
    for i, r in enumerate(res):
        if int(r["score1"]) >= 0:
            out[i] = SswResult(int(r["score1"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]),
                               int(r["read_end1"]))
    return out


def ssw(s1: str, s2: str, open: int, extend: int, matrix: Matrix, engine: Optional[Engine] = None) -> Optional[SswResult]:
    return ssw_batch([(s1, s2)], open, extend, matrix, engine)[0]
