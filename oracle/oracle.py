"""CPU oracle for nanomonsv's supporting-read validation path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module.  The product package (nanomonsv_b200) never does; it fails loudly without its CUDA library.

What is restated here (reference = /root/reference, nanomonsv v0.8.0):
  * parasail.ssw + matrix_create("ACGT", 2, -2)      -> oracle/nmsv_oracle.c (scalar), oracle/nmsv_striped.c (AVX2)
    (call sites count_sread_by_alignment.py:139,148-149; parasail itself is an absent third-party dependency,
    pinned parasail==1.3.4 in the reference Dockerfile:49 -- PARITY UNPINNED against the real binary)
  * ssw_check strand rule and 1-based coordinates      (count_sread_by_alignment.py:137-164)
  * my_seq.reverse_complement                          (my_seq.py:20-26)
  * pyssw.read FASTA id / duplicate semantics          (pyssw.py:25-47)
  * Alignment_counter.initialize / generate_segment_fasta_files[_sbnd] / count_alignment_and_print[_sbnd]
                                                       (count_sread_by_alignment.py:207-397)
  * the streaming group-by of count_sread_by_alignment (count_sread_by_alignment.py:414-448)

Pins: the host-logic restatements above (everything but the parasail.ssw arithmetic) are PINNED on a run of the
reference's own count_sread_by_alignment (tests/golden/stage_reference.json, made by tests/golden/make_golden_stage.py
with stand-in pysam / parasail modules; checked by tests/test_stage_reference.py). The arithmetic is pinned on the
hand-derived and machine-made vectors under tests/golden/ only: PARITY UNPINNED against a real parasail binary.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnmsv_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (oracle/Makefile). Returns the path of the shared library."""
    srcs = [os.path.join(_HERE, f) for f in ("nmsv_oracle.c", "nmsv_striped.c", "nmsv_jump_oracle.c")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libnmsv_oracle.so"])
    return _LIB_PATH


class SswResult(ctypes.Structure):
    """Field names follow the parasail.ssw result object used at count_sread_by_alignment.py:151-159."""
    _fields_ = [("score1", ctypes.c_int32), ("ref_begin1", ctypes.c_int32), ("ref_end1", ctypes.c_int32),
                ("read_begin1", ctypes.c_int32), ("read_end1", ctypes.c_int32)]

    def astuple(self) -> Tuple[int, int, int, int, int]:
        return (self.score1, self.read_begin1, self.read_end1, self.ref_begin1, self.ref_end1)


class JumpResult(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("found", "score", "i1_start", "i1_end", "i2_start", "i2_end", "j1_start",
                                              "j1_end", "j2_start", "j2_end")] + \
               [("n_ops2", ctypes.c_uint32), ("n_ops1", ctypes.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.ora_ssw.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(SswResult)]
        L.ora_ssw.restype = ctypes.c_int
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.ora_ssw_codes.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.POINTER(SswResult)]
        L.ora_ssw_codes.restype = ctypes.c_int
        L.ora_striped_ssw_codes.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(SswResult)]
        L.ora_striped_ssw_codes.restype = ctypes.c_int
        vp = ctypes.c_void_p
        L.ora_ssw_batch.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, vp, vp]
        L.ora_ssw_batch.restype = ctypes.c_int
        L.ora_striped_ssw_batch.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
        L.ora_striped_ssw_batch.restype = ctypes.c_int
        L.ora_have_avx2.restype = ctypes.c_int
        L.ora_sw_jump.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(JumpResult), vp]
        L.ora_sw_jump.restype = ctypes.c_int
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------------------------
# operator level: parasail.ssw
# ----------------------------------------------------------------------------------------------------------------
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq: str) -> np.ndarray:
    """parasail matrix_create("ACGT", ...) mapper: case-insensitive ACGT -> 0..3, every other byte -> 4."""
    return _CODE[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]


def ssw(s1: str, s2: str, gap_open: int = 3, gap_extend: int = 1, match: int = 2, mismatch: int = -2,
        engine: str = "scalar") -> Optional[SswResult]:
    """parasail.ssw(s1, s2, gap_open, gap_extend, matrix_create("ACGT", match, mismatch)); None on saturation."""
    L = lib()
    res = SswResult()
    if engine == "scalar":
        rc = L.ora_ssw(s1.encode("latin-1"), len(s1), s2.encode("latin-1"), len(s2), match, mismatch,
                       gap_open, gap_extend, ctypes.byref(res))
    else:
        c1, c2 = np.ascontiguousarray(encode(s1)), np.ascontiguousarray(encode(s2))
        u8p = ctypes.POINTER(ctypes.c_uint8)
        rc = L.ora_striped_ssw_codes(c1.ctypes.data_as(u8p), len(s1), c2.ctypes.data_as(u8p), len(s2), match,
                                     mismatch, gap_open, gap_extend, 1 if engine == "striped16" else 0,
                                     ctypes.byref(res))
    if rc < 0:
        raise MemoryError("oracle allocation failure")
    return None if rc == 1 else res


RESULT_DTYPE = np.dtype([("score1", "<i4"), ("ref_begin1", "<i4"), ("ref_end1", "<i4"),
                         ("read_begin1", "<i4"), ("read_end1", "<i4")])


def ssw_batch(codes: np.ndarray, offsets: np.ndarray, lens: np.ndarray, tmpl_idx: np.ndarray, read_idx: np.ndarray,
              gap_open: int = 3, gap_extend: int = 1, match: int = 2, mismatch: int = -2,
              engine: str = "scalar", threads: int = 0) -> Tuple[np.ndarray, int]:
    """Batched pair alignment over a packed code buffer. Returns (results[RESULT_DTYPE], threads used)."""
    L = lib()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    tmpl_idx = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
    read_idx = np.ascontiguousarray(read_idx, dtype=np.uint32)
    out = np.zeros(len(tmpl_idx), dtype=RESULT_DTYPE)
    if engine == "scalar":
        rc = L.ora_ssw_batch(codes.ctypes.data, offsets.ctypes.data, lens.ctypes.data, tmpl_idx.ctypes.data,
                             read_idx.ctypes.data, len(tmpl_idx), match, mismatch, gap_open, gap_extend,
                             out.ctypes.data, None)
        used = 1
    else:
        rc = L.ora_striped_ssw_batch(codes.ctypes.data, offsets.ctypes.data, lens.ctypes.data, tmpl_idx.ctypes.data,
                                     read_idx.ctypes.data, len(tmpl_idx), match, mismatch, gap_open, gap_extend,
                                     threads, 1 if engine == "striped16" else 0, out.ctypes.data, None)
        used = rc
    if rc < 0:
        raise MemoryError("oracle allocation failure")
    return out, used


def ssw_pure_python(s1: str, s2: str, gap_open: int = 3, gap_extend: int = 1, match: int = 2,
                    mismatch: int = -2) -> Tuple[int, int, int, int, int]:
    """Tiny independent pure-Python restatement (full matrices); small inputs only.
    Returns (score1, read_begin1, read_end1, ref_begin1, ref_end1)."""

    def sc(a: str, b: str) -> int:
        a, b = a.upper(), b.upper()
        if a not in "ACGT" or b not in "ACGT":
            return 0
        return match if a == b else mismatch

    def one_pass(a: str, b: str) -> Tuple[int, int, int]:
        m, n = len(a), len(b)
        neg = -10 ** 9
        H = [[0] * (n + 1) for _ in range(m + 1)]
        E = [[neg] * (n + 1) for _ in range(m + 1)]
        F = [[neg] * (n + 1) for _ in range(m + 1)]
        for i in range(1, m + 1):
            for j in range(1, n + 1):
                E[i][j] = max(E[i][j - 1] - gap_extend, H[i][j - 1] - gap_open)
                F[i][j] = max(F[i - 1][j] - gap_extend, H[i - 1][j] - gap_open)
                H[i][j] = max(0, H[i - 1][j - 1] + sc(a[i - 1], b[j - 1]), E[i][j], F[i][j])
        best, bi, bj = 0, m - 1, 0
        for j in range(1, n + 1):          # first column that raises the maximum ...
            for i in range(1, m + 1):      # ... smallest row in it
                if H[i][j] > best:
                    best, bi, bj = H[i][j], i - 1, j - 1
        return best, bi, bj

    score, ei, ej = one_pass(s1, s2)
    _, ri, rj = one_pass(s1[:ei + 1][::-1], s2[:ej + 1][::-1])
    return score, ei - ri, ei, ej - rj, ej


# ----------------------------------------------------------------------------------------------------------------
# function level: my_seq.reverse_complement, pyssw.read, ssw_check
# ----------------------------------------------------------------------------------------------------------------
_COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "W": "W", "S": "S", "M": "K", "K": "M", "R": "Y", "Y": "R",
               "B": "V", "V": "B", "D": "H", "H": "D", "N": "N"}


def reverse_complement(seq: str) -> str:
    """my_seq.py:20-26 -- upper-case IUPAC table, every other character (incl. lower case) passes through."""
    return "".join(_COMPLEMENT.get(b, b) for b in reversed(seq))


def fasta_records(path: str) -> Iterable[Tuple[str, str]]:
    """pyssw.py:31-47 -- id = first whitespace token of the header; a record with an empty sequence is dropped
    unless it is the last one in the file."""
    sid, seq = "", ""
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if seq:
                    yield sid, seq
                sid = line.strip()[1:].split()[0]
                seq = ""
            else:
                seq += line.strip()
    yield sid, seq


def ssw_check(template: str, reads: Sequence[Tuple[str, str]], engine: str = "scalar") -> Dict[str, list]:
    """count_sread_by_alignment.py:137-164 with the two FASTA files already parsed: for one template and every
    read align both template strands, keep '+' iff score_f > score_r (ties -> '-'), 1-based coordinates."""
    info: Dict[str, list] = {}
    template_r = reverse_complement(template)
    for rid, rseq in reads:
        res = ssw(template, rseq, engine=engine)
        res_r = ssw(template_r, rseq, engine=engine)
        if res is None or res_r is None:
            raise AttributeError("parasail.ssw returned None (16-bit saturation); the reference would crash here")
        if res.score1 > res_r.score1:
            info[rid] = [res.score1, res.read_begin1 + 1, res.read_end1 + 1, res.ref_begin1 + 1, res.ref_end1 + 1, "+"]
        else:
            info[rid] = [res_r.score1, len(template) - res_r.read_end1, len(template) - res_r.read_begin1,
                         res_r.ref_begin1 + 1, res_r.ref_end1 + 1, "-"]
    return info


# ----------------------------------------------------------------------------------------------------------------
# stage level: Alignment_counter on the sorted seam TSV
# ----------------------------------------------------------------------------------------------------------------
class DictFasta:
    """Duck-typed stand-in for pysam.FastaFile.fetch(chrom, start0, end0) (half-open, clipped at contig end)."""

    def __init__(self, contigs: Dict[str, str]):
        self.contigs = contigs

    def fetch(self, chrom: str, start: int, end: int) -> str:
        return self.contigs[chrom][max(start, 0):max(end, 0)]

    def close(self) -> None:
        pass


def sv_templates(key: str, fasta, L: int = 200) -> Tuple[str, str, str, str, bool]:
    """count_sread_by_alignment.py:207-264: (var1, var2, ref1, ref2, is_short_del_dup) for a canonical SV key."""
    f = key.split(",")
    # :210-213 -- the id replaces the insertion column, then its *length as a string* is stored
    f2 = list(f)
    f2[-2] = "" if f2[-2] == "---" else f2[-1]
    f2[-2] = str(len(f2[-2]))
    span_extra = len(f2[6])  # :219-222 use len() of that string: 1 or 2 characters
    short = False
    if f2[0] == f2[3] and ((f2[2] == "+" and f2[5] == "-") or (f2[2] == "-" and f2[5] == "+")):
        if int(f2[4]) - int(f2[1]) + span_extra < 300:
            short = True
    chr1, pos1, dir1, chr2, pos2, dir2, inseq = f[0], int(f[1]), f[2], f[3], int(f[4]), f[5], f[6]
    inseq = "" if inseq == "---" else inseq
    ref1 = fasta.fetch(chr1, max(pos1 - L - 1, 0), pos1 + L - 1).upper()
    ref2 = fasta.fetch(chr2, max(pos2 - L - 1, 0), pos2 + L - 1).upper()
    if dir1 == "+":
        left = fasta.fetch(chr1, max(pos1 - L - 1, 0), pos1 - 1).upper()
        variant = left + inseq
    else:
        left = reverse_complement(fasta.fetch(chr1, pos1 - 1, pos1 + L - 1).upper())
        variant = left + reverse_complement(inseq)
    if dir2 == "-":
        right = fasta.fetch(chr2, pos2 - 1, pos2 + L - 1).upper()
    else:
        right = reverse_complement(fasta.fetch(chr2, max(pos2 - L - 1, 0), pos2 - 1).upper())
    variant += right
    n = min(2 * L, len(variant))
    return variant[:n], variant[-n:], ref1, ref2, short


def sbnd_templates(key: str, fasta, L: int = 200) -> Tuple[str, str]:
    """count_sread_by_alignment.py:267-283: (var1, ref1) for a single-breakend key; '+' appends the contig twice."""
    chrom, pos, direction, contig, _ = key.split(",")
    pos = int(pos)
    ref1 = fasta.fetch(chrom, max(pos - L - 1, 0), pos + L - 1).upper()
    if direction == "+":
        seg = fasta.fetch(chrom, max(pos - L - 1, 0), pos).upper() + contig
    else:
        seg = reverse_complement(fasta.fetch(chrom, pos - 1, pos + L - 1).upper())
    return seg + contig, ref1


def _mapq(v: str) -> Optional[int]:
    return None if v == "None" else int(v)


def count_one_sv(key: str, rows: Sequence[Tuple[str, str, str, str]], fasta, *, is_sbnd: bool = False,
                 var_read_min_mapq: int = 0, score_ratio_thres: float = 1.2, validate_sequence_length: int = 200,
                 start_pos_thres: float = 0.1, end_pos_thres: float = 0.9, var_ref_margin_thres: int = 20,
                 engine: str = "scalar") -> Tuple[str, List[str]]:
    """One key group of the seam TSV -> (sread_count line, list of sread_info lines).
    rows = [(read id, mapq1, mapq2, sequence)] in file order (count_sread_by_alignment.py:286-397)."""
    mapq: Dict[str, list] = {}
    fasta_reads: List[Tuple[str, str]] = []
    for rid, m1, m2, seq in rows:
        mapq[rid] = [_mapq(m1), _mapq(m2)]
        # the reference round-trips the reads through a FASTA file: the id is cut at the first blank and an
        # empty sequence makes pyssw.read skip the record
        fasta_reads.append((rid.split()[0] if rid.split() else "", seq.strip()))
    reads: List[Tuple[str, str]] = [(i, s) for k, (i, s) in enumerate(fasta_reads) if s or k == len(fasta_reads) - 1]

    if is_sbnd:
        var1, ref1 = sbnd_templates(key, fasta, validate_sequence_length)
        a_var1 = ssw_check(var1, reads, engine)
        a_ref1 = ssw_check(ref1, reads, engine)
        names = list(dict.fromkeys(a_var1.keys()))
        supp = [r for r in names if a_var1[r][0] > score_ratio_thres * len(var1)
                and a_var1[r][1] < start_pos_thres * len(var1) and a_var1[r][2] > end_pos_thres * len(var1)]
        supp = [r for r in supp if a_var1[r][0] >= a_ref1[r][0] + var_ref_margin_thres]
        supp = [r for r in supp if mapq[r][0] is not None and mapq[r][0] >= var_read_min_mapq]
        chrom, pos, direction, contig, sid = key.split(",")
        head = f"{chrom}\t{pos}\t{direction}\t{contig}\t{sid}"
        count_line = f"{head}\t{len(names)}\t{len(supp)}"
        info = [f"{head}\t{r}\t" + "\t".join(str(x) for x in a_var1[r]) for r in supp]
        return count_line, info

    var1, var2, ref1, ref2, short = sv_templates(key, fasta, validate_sequence_length)
    a_var1 = ssw_check(var1, reads, engine)
    a_var2 = ssw_check(var2, reads, engine)   # is_inseq is always True (:211-215), so var2 is always aligned
    if short:
        a_ref1 = ssw_check(ref1, reads, engine)
        a_ref2 = ssw_check(ref2, reads, engine)
    names = list(dict.fromkeys(list(a_var1.keys()) + list(a_var2.keys())))

    def covers(a, tmpl):
        return (a[0] > score_ratio_thres * len(tmpl) and a[1] < start_pos_thres * len(tmpl)
                and a[2] > end_pos_thres * len(tmpl))

    supp = [r for r in names if covers(a_var1[r], var1) or covers(a_var2[r], var2)]
    if short:
        t = var_ref_margin_thres
        supp = [r for r in supp if
                (a_var1[r][0] >= a_ref1[r][0] + t and a_var1[r][0] >= a_ref2[r][0] + t) or
                (a_var2[r][0] >= a_ref1[r][0] + t and a_var2[r][0] >= a_ref2[r][0] + t)]
    supp = [r for r in supp if mapq[r][0] is not None and mapq[r][0] >= var_read_min_mapq
            and mapq[r][1] is not None and mapq[r][1] >= var_read_min_mapq]
    head = "\t".join(key.split(","))
    count_line = f"{head}\t{len(names)}\t{len(supp)}"
    info = []
    for r in supp:
        tail = a_var1[r] + a_var2[r] + ((a_ref1[r] + a_ref2[r]) if short else ["None"] * 12)
        info.append(f"{head}\t{r}\t" + "\t".join(str(x) for x in tail))
    return count_line, info


def count_from_seam(seam_path: str, fasta, output_count_file: str, output_info_file: str, *, is_sbnd: bool = False,
                    **params) -> None:
    """The streaming group-by of count_sread_by_alignment.py:414-425 / :436-448 on an already sorted seam TSV."""
    with open(seam_path) as hin, open(output_count_file, "w") as hc, open(output_info_file, "w") as hi:
        cur_key, rows = None, []

        def flush():
            if cur_key is None:
                return
            count_line, info = count_one_sv(cur_key, rows, fasta, is_sbnd=is_sbnd, **params)
            print(count_line, file=hc)
            for line in info:
                print(line, file=hi)

        for line in hin:
            key, rid, m1, m2, seq = line.rstrip("\n").split("\t")
            if is_sbnd:
                m2 = "None"
            if key != cur_key:
                flush()
                cur_key, rows = key, []
            rows.append((rid, m1, m2, seq))
        flush()


# ----------------------------------------------------------------------------------------------------------------
# "next" row: smith_waterman.sw_jump (reference nanomonsv/smith_waterman.py:5-127)
# ----------------------------------------------------------------------------------------------------------------
def jump_strings(contig: str, region1: str, region2: str, res, ops: np.ndarray):
    """Rebuild the two alignment strings of sw_jump's return value (:61-125) from the traceback operations."""
    c2, g2 = [], []
    i, j = res["i2_end"], res["j2_end"]
    for p in ops[:res["n_ops2"]]:
        if p == 1:
            c2.append(contig[i - 1]); g2.append(region2[j - 1]); i -= 1; j -= 1
        elif p == 2:
            c2.append(contig[i - 1]); g2.append("-"); i -= 1
        elif p == 3:
            c2.append("-"); g2.append(region2[j - 1]); j -= 1
        elif p == 4:
            c2.append(contig[i - 1]); g2.append(region2[j - 1])
    c1, g1 = [], []
    i, j = res["i1_end"], res["j1_end"]
    for p in ops[res["n_ops2"]:res["n_ops2"] + res["n_ops1"]]:
        if p == 1:
            c1.append(contig[i - 1]); g1.append(region1[j - 1]); i -= 1; j -= 1
        elif p == 2:
            c1.append(contig[i - 1]); g1.append("-"); i -= 1
        elif p == 3:
            c1.append("-"); g1.append(region1[j - 1]); j -= 1
    return "".join(reversed(c1)) + " " + "".join(reversed(c2)), "".join(reversed(g1)) + " " + "".join(reversed(g2))


def sw_jump(contig: str, region1_seq: str, region2_seq: str, match_score: int = 1, mismatch_penalty: int = 2,
            gap_cost: int = 2, jump_cost: int = 2, minimum_score: int = 10):
    """smith_waterman.sw_jump with the same return value (None or the 6-tuple), computed by the C oracle."""
    L = lib()
    c = np.frombuffer(contig.encode("latin-1"), dtype=np.uint8)
    a = np.frombuffer(region1_seq.encode("latin-1"), dtype=np.uint8)
    b = np.frombuffer(region2_seq.encode("latin-1"), dtype=np.uint8)
    with np.errstate(divide="ignore"):
        logtab = np.log(np.arange(len(c) + 1))
    res = JumpResult()
    ops = np.zeros(2 * len(c) + len(a) + len(b) + 4, dtype=np.uint8)
    rc = L.ora_sw_jump(c.ctypes.data, len(c), a.ctypes.data, len(a), b.ctypes.data, len(b), match_score,
                       mismatch_penalty, gap_cost, jump_cost, minimum_score, logtab.ctypes.data, ctypes.byref(res),
                       ops.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle allocation failure")
    if not res.found:
        return None
    d = {n: getattr(res, n) for n, _ in JumpResult._fields_}
    cm, rm = jump_strings(contig, region1_seq, region2_seq, d, ops)
    return (res.score, (res.i1_start, res.i1_end, res.i2_start, res.i2_end), (res.j1_start, res.j1_end),
            (res.j2_start, res.j2_end), cm, rm)
