"""GPU-backed form of one step of the reference's final filtering: `Sv_filterer.filter_sv_insertion_match`
(nanomonsv/post_proc.py:84-134) as driven by `apply_filters` (:226-232) -- a non-insertion SV is dropped as
"Duplicate_with_insertion" when the sequence across its junction aligns end to end, on either strand, inside the sequence
across a nearby insertion. The reference runs two `parasail.ssw(.., 3, 1, matrix_create("ACGT", 2, -2))` per examined pair, one
pair at a time inside an O(N^2) loop over the SV list; here the alignments of all examined pairs
are one GPU batch.

The loop is order dependent -- a pair is skipped when one of its members has been filtered by an earlier pair -- but what an
examined pair decides depends on its two sequences alone. So: pass 1 walks the pairs in the reference's order and builds the
sequences of every pair that passes the size and position tests (:87-91), pass 2 aligns them all, pass 3 walks the pairs again
in the same order, skips as the reference would, and applies the verdicts. `sv` objects are the reference's `Sv` (or anything
with chr1 pos1 dir1 chr2 pos2 dir2 inseq filter)."""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple

from . import ssw as _ssw
from .engine import Engine
from .my_seq import reverse_complement


def insertion_match_segments(sv, ins, reference_h, min_indel_size=50, bp_dist_margin=30, validate_seg_len=100) -> Optional[Tuple[str, str]]:
    """(sv_seg, ins_seg) of post_proc.py:93-116, or None when the reference returns before aligning (:87-91)."""
    if len(sv.inseq) >= min_indel_size or len(ins.inseq) < min_indel_size:
        return None
    if not (sv.chr1 == ins.chr1 and abs(sv.pos1 - ins.pos1) <= bp_dist_margin) and \
            not (sv.chr2 == ins.chr2 and abs(sv.pos2 - ins.pos2) <= bp_dist_margin):
        return None
    L, M = validate_seg_len, bp_dist_margin
    ins_seg = reference_h.fetch(ins.chr1, max(ins.pos1 - L - M - 1, 0), ins.pos1 - 1).upper()
    ins_seg = ins_seg + ins.inseq
    ins_seg = ins_seg + reference_h.fetch(ins.chr1, ins.pos2 - 1, ins.pos2 + L + M - 1).upper()
    if sv.dir1 == '+':
        tseq = reference_h.fetch(sv.chr1, max(sv.pos1 - L - 1, 0), sv.pos1 - 1).upper()
    else:
        tseq = reverse_complement(reference_h.fetch(sv.chr1, sv.pos1 - 1, sv.pos1 + L - 1).upper())
    sv_seg = tseq + (sv.inseq if sv.dir1 == '+' else reverse_complement(sv.inseq))
    if sv.dir2 == '-':
        tseq = reference_h.fetch(sv.chr2, sv.pos2 - 1, sv.pos2 + L - 1).upper()
    else:
        tseq = reverse_complement(reference_h.fetch(sv.chr2, max(sv.pos2 - L - 1, 0), sv.pos2 - 1).upper())
    return sv_seg + tseq, ins_seg


def insertion_match_verdict(res, res_r, sv_seg_len: int) -> bool:
    """post_proc.py:122-133: the better strand (ties: reverse) must cover the junction sequence end to end at > 0.75 of a
    perfect score per aligned base of the insertion sequence."""
    r = res if res.score1 > res_r.score1 else res_r
    match_ratio = float(r.score1) / (2 * (r.ref_end1 - r.ref_begin1 + 1))
    return match_ratio > 0.75 and r.read_begin1 < 0.1 * sv_seg_len and r.read_end1 > 0.9 * sv_seg_len


def _ordered_pairs(sv_list: Sequence, min_indel_size: int):
    """(sv, ins) in the order and under the size test of apply_filters (:227-232)."""
    for sv1, sv2 in itertools.combinations(sv_list, 2):
        if len(sv1.inseq) < min_indel_size and len(sv2.inseq) >= min_indel_size:
            yield sv1, sv2
        elif len(sv2.inseq) < min_indel_size and len(sv1.inseq) >= min_indel_size:
            yield sv2, sv1


def apply_insertion_match_filter(sv_list: Sequence, reference_h, min_indel_size=50, bp_dist_margin=30, validate_seg_len=100,
                                 filter_item="Duplicate_with_insertion", engine: Optional[Engine] = None) -> int:
    """The `filter_sv_insertion_match` stage of Sv_filterer.apply_filters on `sv_list` (filters appended in place like the
    reference does). Returns the number of alignments that ran on the GPU."""
    examined: List[Tuple[int, str, str]] = []          # (index into the pair order, sv_seg, ins_seg)
    order = list(_ordered_pairs(sv_list, min_indel_size))
    for k, (sv, ins) in enumerate(order):
        segs = insertion_match_segments(sv, ins, reference_h, min_indel_size, bp_dist_margin, validate_seg_len)
        if segs is not None:
            examined.append((k, segs[0], segs[1]))
    pairs = []
    for _, sv_seg, ins_seg in examined:
        pairs.append((sv_seg, ins_seg))
        pairs.append((reverse_complement(sv_seg), ins_seg))
    results = _ssw.ssw_batch(pairs, 3, 1, _ssw.matrix_create("ACGT", 2, -2), engine=engine) if pairs else []
    verdict = {}
    for j, (k, sv_seg, _) in enumerate(examined):
        res, res_r = results[2 * j], results[2 * j + 1]
        if res is None or res_r is None:               # the reference would fail on `.score1` of None here
            raise ValueError("filter_sv_insertion_match: an alignment score is not representable in 16 bits")
        verdict[k] = insertion_match_verdict(res, res_r, len(sv_seg))
    for k, (sv, ins) in enumerate(order):
        if len(sv.filter) > 0 or len(ins.filter) > 0:  # :228 -- the state left by the pairs before this one
            continue
        if verdict.get(k, False):
            sv.filter.append(filter_item)
    return len(pairs)
