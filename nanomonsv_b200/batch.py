"""Packing of SV candidates, their templates and their reads into the device batch layout (include/nmsv.h):
one uint8 code buffer + an nmsv_sv table + an nmsv_read table."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from .engine import AsciiSlice, SeqPack
from .my_seq import needs_explicit_revcomp, reverse_complement


def integer_bounds(length: int, score_ratio_thres: float, start_pos_thres: float, end_pos_thres: float):
    """The three float comparisons of reference count_sread_by_alignment.py:328-330 as integer bounds
    (score_min, qstart_max, qend_min):
        score  > r*len  <=>  score  >= floor(r*len) + 1
        qstart < s*len  <=>  qstart <= ceil(s*len) - 1
        qend   > e*len  <=>  qend   >= floor(e*len) + 1
    The products are Python doubles exactly as in the reference (e.g. 1.8*401 = 721.8000000000001)."""
    return (math.floor(score_ratio_thres * length) + 1,
            math.ceil(start_pos_thres * length) - 1,
            math.floor(end_pos_thres * length) + 1)


class BatchBuilder:
    """Accumulates SVs and reads; finalize() returns (SeqPack, svs[SV_DTYPE], reads[READ_DTYPE])."""

    def __init__(self, score_ratio_thres: float = 1.2, start_pos_thres: float = 0.1, end_pos_thres: float = 0.9,
                 var_ref_margin_thres: int = 20, var_read_min_mapq: int = 0):
        self.score_ratio_thres = score_ratio_thres
        self.start_pos_thres = start_pos_thres
        self.end_pos_thres = end_pos_thres
        self.var_ref_margin_thres = var_ref_margin_thres
        self.var_read_min_mapq = var_read_min_mapq
        self.pack = SeqPack()
        self._svs: List[tuple] = []
        self._reads: List[tuple] = []
        self._open_sv: Optional[int] = None
        self.bases = 0

    def add_sv(self, templates: Sequence[Optional[str]], is_sbnd: bool, short: bool) -> int:
        """templates = [var1, var2, ref1, ref2] (None = slot absent). Reads added next belong to this SV."""
        self._close_sv()
        tmpl = [_lib.NMSV_NONE] * 4
        tmpl_rc = [_lib.NMSV_NONE] * 4
        seen: Dict[str, Tuple[int, int]] = {}
        for s, t in enumerate(templates):
            if t is None:
                continue
            if not t:
                raise ValueError("empty template")
            if t not in seen:   # identical templates of one SV (var1 == var2 without an insertion) share work
                rc = self.pack.add(reverse_complement(t)) if needs_explicit_revcomp(t) else _lib.NMSV_NONE
                seen[t] = (self.pack.add(t), rc)
            tmpl[s], tmpl_rc[s] = seen[t]
        bounds = [(0, 0, 0), (0, 0, 0)]
        for v in (0, 1):
            if templates[v] is not None:
                bounds[v] = integer_bounds(len(templates[v]), self.score_ratio_thres, self.start_pos_thres,
                                           self.end_pos_thres)
        flags = (_lib.SV_SBND if is_sbnd else 0) | (_lib.SV_SHORT_DEL_DUP if short else 0)
        self._svs.append([tmpl, tmpl_rc, len(self._reads), len(self._reads), [b[0] for b in bounds],
                          [b[1] for b in bounds], [b[2] for b in bounds], self.var_ref_margin_thres,
                          self.var_read_min_mapq, flags, 0])
        self._open_sv = len(self._svs) - 1
        return self._open_sv

    def add_read(self, seq: Union[str, np.ndarray], mapq1: Optional[int], mapq2: Optional[int]) -> int:
        assert self._open_sv is not None, "add_sv first"
        if isinstance(seq, str):
            idx = self.pack.add(seq)
        elif isinstance(seq, AsciiSlice):
            idx = self.pack.add_ascii_slice(seq)
        else:
            idx = self.pack.add_codes(seq)
        self.bases += self.pack.length(idx)
        self._reads.append((idx, self._open_sv, _lib.MAPQ_NONE if mapq1 is None else int(mapq1),
                            _lib.MAPQ_NONE if mapq2 is None else int(mapq2)))
        return len(self._reads) - 1

    def add_read_absent(self, length: int, mapq1: Optional[int], mapq2: Optional[int]) -> int:
        """A read that is only counted (its mapq test fails, so it can never support): its sequence is not packed."""
        assert self._open_sv is not None, "add_sv first"
        idx = self.pack.add_absent(length)
        self._reads.append((idx, self._open_sv, _lib.MAPQ_NONE if mapq1 is None else int(mapq1),
                            _lib.MAPQ_NONE if mapq2 is None else int(mapq2)))
        return len(self._reads) - 1

    def read_seq_index(self, read_row: int) -> int:
        return self._reads[read_row][0]

    def add_read_shared(self, seq_index: int, mapq1: Optional[int], mapq2: Optional[int]) -> int:
        """A read whose sequence is already in the pack (the same read serves several SV candidates, reference
        count_sread_by_alignment.py:109-111): only a row of the read table is added."""
        assert self._open_sv is not None, "add_sv first"
        self.bases += self.pack.length(seq_index)
        self._reads.append((seq_index, self._open_sv, _lib.MAPQ_NONE if mapq1 is None else int(mapq1),
                            _lib.MAPQ_NONE if mapq2 is None else int(mapq2)))
        return len(self._reads) - 1

    def _close_sv(self):
        if self._open_sv is not None:
            self._svs[self._open_sv][3] = len(self._reads)
            self._open_sv = None

    @property
    def n_svs(self) -> int:
        return len(self._svs)

    def finalize(self) -> Tuple[SeqPack, np.ndarray, np.ndarray]:
        self._close_sv()
        svs = np.zeros(len(self._svs), dtype=_lib.SV_DTYPE)
        for i, row in enumerate(self._svs):
            svs[i] = tuple(row)
        reads = np.array(self._reads, dtype=_lib.READ_DTYPE) if self._reads else np.zeros(0, dtype=_lib.READ_DTYPE)
        self.pack.finalize()
        return self.pack, svs, reads
