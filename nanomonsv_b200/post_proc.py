"""Consumer side of the supporting-read info file: nanomonsv/post_proc.py:306-340 (`proc_sread_info_file`).

The reference turns every supporting-read line of `*.sread_info.txt` whose SV made it into the result file into
`key <tab> read id <tab> breakpoint position on the read <tab> strand` (`*.supporting_read.txt`). The position is a
linear interpolation along the better of the two variant-segment alignments -- it consumes exactly the 1-based
(qstart, qend, tstart, tend) tuples that the CUDA path produces, which makes it the downstream check of those
coordinates.

Pinned on the reference itself: tests/golden/proc_sread_info_vectors.json holds outputs of the reference's own function
(tests/golden/make_golden_postproc.py). Behaviour kept as written there:
  * the first variant segment wins ties (score1 >= score2);
  * the second segment's position is shifted by len(inserted-sequence FIELD) + 1 -- for "---" that is 3 + 1;
  * any strand other than '+' takes the '-' formula; the result is truncated toward zero by int();
  * a zero-length template span (qend == qstart) raises ZeroDivisionError, as in the reference.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Set, Tuple

KEY_COLUMNS = 8


def called_keys(sv_result_file: str) -> Set[str]:
    """Keys (first eight columns) of the SVs in a result file; the first line is a header."""
    keys: Set[str] = set()
    with open(sv_result_file) as hin:
        hin.readline()
        for line in hin:
            keys.add("\t".join(line.rstrip("\t").split("\t")[:KEY_COLUMNS]))
    return keys


def _interpolate(num: int, span: int, rest: int, origin: int) -> float:
    # (float(num) / float(span)) * rest + origin, evaluated in this order with Python doubles (post_proc.py:327-338)
    return (float(num) / float(span)) * rest + origin


def breakpoint_on_read(fields: Sequence[str], validate_sequence_length: int = 200) -> Tuple[int, str]:
    """(position, strand) of the breakpoint on the read for one info line split into fields (post_proc.py:322-338)."""
    s1, qs1, qe1, ts1, te1 = (int(x) for x in fields[9:14])
    s2, qs2, qe2, ts2, te2 = (int(x) for x in fields[15:20])
    ins_len = len(fields[6])
    L = validate_sequence_length
    if s1 >= s2:
        strand = fields[14]
        if strand == "+":
            pos = _interpolate(te1 - ts1, qe1 - qs1, L - qs1, ts1)
        else:
            pos = _interpolate(ts1 - te1, qe1 - qs1, L - qs1, te1)
    else:
        strand = fields[20]
        if strand == "+":
            pos = _interpolate(te2 - ts2, qe2 - qs2, L - qs2, ts2) - ins_len - 1
        else:
            pos = _interpolate(ts2 - te2, qe2 - qs2, L - qs2, te2) + ins_len + 1
    return int(pos), strand


def supporting_read_lines(info_lines: Iterable[str], keys: Set[str], validate_sequence_length: int = 200) -> Iterator[str]:
    for line in info_lines:
        fields = line.rstrip("\n").split("\t")
        key = "\t".join(fields[:KEY_COLUMNS])
        if key not in keys:
            continue
        pos, strand = breakpoint_on_read(fields, validate_sequence_length)
        yield f"{key}\t{fields[8]}\t{pos}\t{strand}"


def proc_sread_info_file(tumor_sread_info_file, sv_result_file, output_file, validate_sequence_length=200):
    """Same signature and output as the reference (post_proc.py:306)."""
    keys = called_keys(sv_result_file)
    with open(tumor_sread_info_file) as hin, open(output_file, "w") as hout:
        for out in supporting_read_lines(hin, keys, validate_sequence_length):
            print(out, file=hout)
