"""GPU-backed drop-in for nanomonsv/count_sread_by_alignment.py (reference v0.8.0).

Same public names, argument meaning, output files and formats as the reference module, so `nanomonsv get` /
`nanomonsv validate` (run.py:156-184, :501-511) can import this module instead:

    gather_local_read_for_realignment   (:27-134)   host, pysam -- unchanged behaviour, writes the seam TSV
    ssw_check                           (:137-164)  one template FASTA x reads FASTA -> dict, on the GPU
    Alignment_counter                   (:167-397)  template construction + decision rules; alignment on the GPU
    count_sread_by_alignment            (:401-451)  driver

What changes is only WHERE the arithmetic runs: instead of 2-4 parasail.ssw calls per read, whole groups of SV
candidates are packed into byte-coded batches and handed to libnmsv_sw.so (nmsv_validate_batch), which returns
the per-SV counts and one record per supporting read.  The reference's float thresholds are converted to integer
bounds here, with the same Python double arithmetic, before they go to the device.

Quirks of the reference that change outputs are reproduced on purpose: var2 is always
aligned; is_short_del_dup uses span + len(str(len(id))) < 300; the '+' single-breakend template carries the
contig twice; strand ties go to '-'.
"""
from __future__ import annotations

import os
import random
import subprocess
from dataclasses import dataclass, field
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .batch import BatchBuilder, integer_bounds  # noqa: F401
from .engine import DEFAULT_SCORING, AsciiSlice, Engine, SeqPack, default_engine
from .my_seq import needs_explicit_revcomp, reverse_complement

# maximum number of read bases packed into one device batch (host memory / H2D granularity, not a GPU limit)
# Measured on the bench's seam files (one Python process, one B200, candidates/s through count_sread_from_seam, batch
# size x calls in flight): 16 MB x2 932, 32 MB x2 1131, 64 MB x2 1244, 16 MB x3 1238, 32 MB x3 1260, 64 MB x3 1258,
# 16 MB x4 1308, 32 MB x4 1364, 64 MB x4 1415 (device resident: 1644) -- the host is ahead of the GPU, what counts is that
# a launch is not all tail and that the end of one launch is filled by the next.
BATCH_BASES = int(os.environ.get("NMSV_BATCH_BASES", str(64 << 20)))
# GPU calls in flight per Alignment_counter (each on a context of its own): the end of a launch runs at low occupancy,
# the next calls fill it
INFLIGHT = int(os.environ.get("NMSV_STAGE_INFLIGHT", "4"))


# ------------------------------------------------------------------------------------------------------------------
# read gathering (host, pysam) -- reference :12-134
# ------------------------------------------------------------------------------------------------------------------
def bam_subsample_fetch(alignment_h, tchr, tstart, tend, max_read_num=500):
    """Reads overlapping a window; above max_read_num an (unseeded, like the reference) random subset."""
    total = sum(1 for _ in alignment_h.fetch(tchr, tstart, tend))
    keep = set(random.sample(range(total), min(max_read_num, total)))
    for i, read in enumerate(alignment_h.fetch(tchr, tstart, tend)):
        if i in keep:
            yield read


def _get_alignment_object(alignment_file, reference_fasta):
    import pysam  # lazy: the GPU path itself does not need pysam
    if alignment_file.endswith(".cram"):
        return pysam.AlignmentFile(alignment_file, "rc", reference_filename=reference_fasta)
    return pysam.AlignmentFile(alignment_file, "rb")


def _window_pass(alignment_h, sv_file, sbnd_file, validate_sequence_length):
    """First pass of the gathering (reference :31-88): which reads lie in a +-100 bp window of which breakpoint, and
    with which mapping quality (the last record of a read name in a window wins, whatever its flags)."""
    rname2key: Dict[str, list] = {}
    key2rname2mapq: Dict[str, dict] = {}
    with open(sv_file) as hin:
        for line in hin:
            if line.startswith("#") or line.startswith("Chr_1"):
                continue
            F = line.rstrip("\n").split("\t")
            key = f"{F[0]},{int(F[1])},{F[2]},{F[3]},{int(F[4])},{F[5]},{F[6]},{F[7]}"
            per_key = key2rname2mapq.setdefault(key, {})
            for which, (c, p) in enumerate(((F[0], int(F[1])), (F[3], int(F[4])))):
                for read in bam_subsample_fetch(alignment_h, c, max(p - 100, 0), p + 100):
                    rname2key.setdefault(read.qname, []).append(key)
                    per_key.setdefault(read.qname, [None, None])[which] = read.mapping_quality
    for rname in rname2key:
        rname2key[rname] = list(set(rname2key[rname]))

    rname2key_sbnd: Dict[str, list] = {}
    key2rname2mapq_sbnd: Dict[str, dict] = {}
    if sbnd_file is not None:
        with open(sbnd_file) as hin:
            for line in hin:
                if line.startswith("#") or line.startswith("Chr_1"):
                    continue
                F = line.rstrip("\n").split("\t")
                pos = int(F[1])
                key = f"{F[0]},{pos},{F[2]},{F[3][:validate_sequence_length]},{F[4]}"
                per_key = key2rname2mapq_sbnd.setdefault(key, {})
                for read in bam_subsample_fetch(alignment_h, F[0], max(pos - 100, 0), pos + 100):
                    rname2key_sbnd.setdefault(read.qname, []).append(key)
                    per_key[read.qname] = read.mapping_quality
        for rname in rname2key_sbnd:
            rname2key_sbnd[rname] = list(set(rname2key_sbnd[rname]))

    return rname2key, key2rname2mapq, rname2key_sbnd, key2rname2mapq_sbnd


def gather_local_read_for_realignment(sv_file, alignment_file, output_file, reference_fasta,
                                      sbnd_file=None, output_file_sbnd=None, validate_sequence_length=200,
                                      check_read_max_num=500, sort_option=None):
    """Writes `key \\t qname \\t mapq1 \\t mapq2 \\t sequence` for every primary read near a breakpoint, sorted by
    key with GNU sort (reference :27-134). check_read_max_num is accepted and, as in the reference, not forwarded."""
    alignment_h = _get_alignment_object(alignment_file, reference_fasta)
    rname2key, key2rname2mapq, rname2key_sbnd, key2rname2mapq_sbnd = _window_pass(alignment_h, sv_file, sbnd_file,
                                                                                  validate_sequence_length)

    hout = open(output_file + ".tmp.unsorted", "w")
    hout_sbnd = open(output_file_sbnd + ".tmp.unsorted", "w") if sbnd_file is not None else None
    for read in alignment_h.fetch():
        if read.is_supplementary or read.is_secondary or read.is_duplicate:
            continue
        in_main = read.qname in rname2key
        in_sbnd = sbnd_file is not None and read.qname in rname2key_sbnd
        if not (in_main or in_sbnd):
            continue
        read_seq = read.query_sequence
        if read.is_reverse:
            read_seq = reverse_complement(read_seq)
        if in_main:
            for key in rname2key[read.qname]:
                mapq1, mapq2 = key2rname2mapq[key][read.qname]
                print(f"{key}\t{read.qname}\t{mapq1}\t{mapq2}\t{read_seq}", file=hout)
        if in_sbnd:
            for key in rname2key_sbnd[read.qname]:
                print(f"{key}\t{read.qname}\t{key2rname2mapq_sbnd[key][read.qname]}\tNone\t{read_seq}", file=hout_sbnd)
    hout.close()
    if hout_sbnd is not None:
        hout_sbnd.close()

    def gnu_sort(path):
        with open(path, "w") as h:
            subprocess.call(["sort", "-k1,1"] + sort_option.split(" ") + [path + ".tmp.unsorted"], stdout=h)
        os.remove(path + ".tmp.unsorted")

    gnu_sort(output_file)
    if sbnd_file is not None:
        gnu_sort(output_file_sbnd)
    alignment_h.close()


def gather_reads_direct(sv_file, alignment_file, reference_fasta, sbnd_file=None, validate_sequence_length=200,
                        sort_option=None):
    """The gathering without the seam file: the same two passes over the alignment file as
    gather_local_read_for_realignment, but the reads stay in memory -- no 'whole read per (read, key)' TSV, no GNU sort
    of it, no re-parse -- and one str object per read is handed to every SV candidate the read serves, so that a batch
    packs and uploads it once. Returns (groups, groups_sbnd): lists of (key, [(qname, mapq1, mapq2, sequence), ...]).

    Order. The reference streams the key-sorted TSV, so its count file lists the keys in the collation order of GNU
    `sort -k1,1` under the caller's locale: the (small) list of keys is therefore sorted by the same `sort` with the same
    options. Within a key the TSV lines are ordered by the rest of the line; here by (read name, mapq1, mapq2, sequence)
    as text in code-point order, which is that order under LC_ALL=C. Only two things see the within-key order: which
    of two primary records with the SAME read name wins (the reference keeps the last, :162) and the order of a
    candidate's supporting-read lines, which the reference itself leaves to set iteration (:325)."""
    alignment_h = _get_alignment_object(alignment_file, reference_fasta)
    rname2key, key2rname2mapq, rname2key_sbnd, key2rname2mapq_sbnd = _window_pass(alignment_h, sv_file, sbnd_file,
                                                                                  validate_sequence_length)
    per_key: Dict[str, list] = {}
    per_key_sbnd: Dict[str, list] = {}
    for read in alignment_h.fetch():
        if read.is_supplementary or read.is_secondary or read.is_duplicate:
            continue
        in_main = read.qname in rname2key
        in_sbnd = sbnd_file is not None and read.qname in rname2key_sbnd
        if not (in_main or in_sbnd):
            continue
        read_seq = read.query_sequence
        if read.is_reverse:
            read_seq = reverse_complement(read_seq)
        if in_main:
            for key in rname2key[read.qname]:
                mapq1, mapq2 = key2rname2mapq[key][read.qname]
                per_key.setdefault(key, []).append((read.qname, mapq1, mapq2, read_seq))
        if in_sbnd:
            for key in rname2key_sbnd[read.qname]:
                per_key_sbnd.setdefault(key, []).append((read.qname, key2rname2mapq_sbnd[key][read.qname], None, read_seq))
    alignment_h.close()

    def in_sort_order(groups: Dict[str, list]) -> list:
        if not groups:
            return []
        proc = subprocess.run(["sort", "-k1,1"] + (sort_option.split(" ") if sort_option else []),
                              input="".join(k + "\n" for k in groups), capture_output=True, text=True, check=True)
        keys = proc.stdout.splitlines()
        assert len(keys) == len(groups)
        return [(k, sorted(groups[k], key=lambda r: (r[0], str(r[1]), str(r[2]), r[3]))) for k in keys]

    return in_sort_order(per_key), in_sort_order(per_key_sbnd)


def _count_from_groups(groups, counter: "Alignment_counter", is_sbnd: bool) -> None:
    for key, rows in groups:
        counter.initialize(key, is_sbnd=is_sbnd)
        for qname, mapq1, mapq2, seq in rows:
            counter.add_long_read_seq(qname, seq, str(mapq1), "None" if is_sbnd else str(mapq2))
        counter.count_alignment_and_print_sbnd() if is_sbnd else counter.count_alignment_and_print()
    counter.close()


# ------------------------------------------------------------------------------------------------------------------
# small parsers shared by ssw_check and the counter
# ------------------------------------------------------------------------------------------------------------------
def read_fasta(path: str):
    """(id, sequence) records with the semantics of pyssw.read (pyssw.py:31-47): id = first token of the header,
    lines concatenated and stripped, a record with an empty sequence is only emitted when it is the last one."""
    sid, parts, have = "", [], False
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                seq = "".join(parts)
                if seq:
                    yield sid, seq
                sid, parts, have = line.strip()[1:].split()[0], [], True
            else:
                parts.append(line.strip())
    if have or parts:
        yield sid, "".join(parts)


def _aln_list(a) -> list:
    """nmsv_aln -> the 6-element list of reference ssw_check (:162)."""
    return [int(a["score"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"]),
            "+" if a["strand"] > 0 else "-"]


def ssw_check(query, target, use_ssw=False, engine: Optional[Engine] = None):
    """Reference ssw_check (:137-164): the template(s) in FASTA `query` against every read of FASTA `target`, both
    strands, keep '+' only if strictly better; returns {read id: [score, qstart, qend, tstart, tend, strand]}.
    `use_ssw` is accepted and ignored, as in the reference."""
    eng = engine or default_engine()
    info: Dict[str, list] = {}
    reads = list(read_fasta(target))
    for _, tseq in read_fasta(query):
        if not reads:
            continue
        pack = SeqPack()
        t = pack.add(tseq)
        rc = pack.add(reverse_complement(tseq)) if needs_explicit_revcomp(tseq) else _lib.NMSV_NONE
        ridx = [pack.add(rseq) for _, rseq in reads]
        res = eng.ssw_check_batch(pack, [t] * len(ridx), ridx, [rc] * len(ridx))
        for (rid, _), a in zip(reads, res):
            info[rid] = _aln_list(a)
    return info


# ------------------------------------------------------------------------------------------------------------------
# template construction and decision thresholds -- reference :207-283, :327-346
# ------------------------------------------------------------------------------------------------------------------
def canonical_templates(key: str, fasta, L: int = 200) -> Tuple[str, str, str, str, bool]:
    """(variant_segment_1, variant_segment_2, reference_segment_1, reference_segment_2, is_short_del_dup)."""
    chr1, pos1, dir1, chr2, pos2, dir2, inseq, sv_id = key.split(",")
    pos1, pos2 = int(pos1), int(pos2)
    # quirk (:210-223): the column compared against 300 is len(str(len(id))) (or len("0") when there is no
    # insertion), not the insertion length
    span_extra = len(str(len("" if inseq == "---" else sv_id)))
    short = (chr1 == chr2 and (dir1, dir2) in (("+", "-"), ("-", "+")) and pos2 - pos1 + span_extra < 300)
    ins = "" if inseq == "---" else inseq

    def ref(c, a, b):
        return fasta.fetch(c, max(a, 0), b).upper()

    ref1 = ref(chr1, pos1 - L - 1, pos1 + L - 1)
    ref2 = ref(chr2, pos2 - L - 1, pos2 + L - 1)
    if dir1 == "+":
        variant = ref(chr1, pos1 - L - 1, pos1 - 1) + ins
    else:
        variant = reverse_complement(fasta.fetch(chr1, pos1 - 1, pos1 + L - 1).upper()) + reverse_complement(ins)
    if dir2 == "-":
        variant += fasta.fetch(chr2, pos2 - 1, pos2 + L - 1).upper()
    else:
        variant += reverse_complement(ref(chr2, pos2 - L - 1, pos2 - 1))
    n = min(2 * L, len(variant))
    return variant[:n], variant[len(variant) - n:], ref1, ref2, short


def sbnd_templates(key: str, fasta, L: int = 200) -> Tuple[str, str]:
    """(variant_segment_1, reference_segment_1) of a single-breakend key `chr,pos,dir,contig,id` (:267-283)."""
    chrom, pos, direction, contig, _ = key.split(",")
    pos = int(pos)
    ref1 = fasta.fetch(chrom, max(pos - L - 1, 0), pos + L - 1).upper()
    if direction == "+":
        head = fasta.fetch(chrom, max(pos - L - 1, 0), pos).upper() + contig   # the contig is appended here ...
    else:
        head = reverse_complement(fasta.fetch(chrom, pos - 1, pos + L - 1).upper())
    return head + contig, ref1                                                 # ... and again here (:279,:283)


@dataclass
class SvJob:
    key: str
    is_sbnd: bool
    templates: List[Optional[str]]           # var1, var2, ref1, ref2
    short: bool
    read_ids: List[str] = field(default_factory=list)
    read_seqs: Dict[str, str] = field(default_factory=dict)     # id -> sequence (last record wins, like the dict in :162)
    read_mapq: Dict[str, Tuple[Optional[int], Optional[int]]] = field(default_factory=dict)
    bases: int = 0

    def read_id(self, ordinal: int) -> str:
        return self.read_ids[ordinal]


class _ScannedSv:
    """Per-SV bookkeeping of a batch assembled from nmsv_seam_scan tables: what _drain needs to write the lines."""
    __slots__ = ("key", "is_sbnd", "short", "data", "sreads", "first")

    def __init__(self, key, is_sbnd, short, data, sreads, first):
        self.key, self.is_sbnd, self.short, self.data, self.sreads, self.first = key, is_sbnd, short, data, sreads, first

    def read_id(self, ordinal: int) -> str:
        r = self.sreads[self.first + ordinal]
        return self.data[int(r["id_off"]):int(r["id_off"]) + int(r["id_len"])].decode()


def _mapq(v) -> Optional[int]:
    if v is None or v == "None":
        return None
    return int(v)


class Alignment_counter(object):
    """Same constructor and methods as the reference class (:167-397). SV candidates are queued and sent to the GPU in
    batches; call flush() (or let count_sread_by_alignment do it) to write everything that is pending."""

    def __init__(self, output_count_file, output_alignment_info_file, reference_fasta,
                 var_read_min_mapq, score_ratio_thres, use_ssw_lib, debug, validate_sequence_length=200,
                 start_pos_thres=0.1, end_pos_thres=0.9, var_ref_margin_thres=20, engine: Optional[Engine] = None,
                 scoring=DEFAULT_SCORING):
        self.temp_key = None
        self.hout_count = open(output_count_file, "w")
        self.hout_ainfo = open(output_alignment_info_file, "w")
        if isinstance(reference_fasta, str):
            import pysam  # lazy
            self.reference_fasta_h = pysam.FastaFile(reference_fasta)
        else:
            self.reference_fasta_h = reference_fasta      # any object with fetch(chrom, start0, end0)
        self.validate_sequence_length = validate_sequence_length
        self.score_ratio_thres = score_ratio_thres
        self.start_pos_thres = start_pos_thres
        self.end_pos_thres = end_pos_thres
        self.var_ref_margin_thres = var_ref_margin_thres
        self.var_read_min_mapq = var_read_min_mapq
        self.use_ssw_lib = use_ssw_lib          # accepted, ignored (reference :137)
        self.debug = debug
        self.scoring = scoring
        self._engine = engine
        self._job: Optional[SvJob] = None
        self._queue: List[SvJob] = []
        self._queued_bases = 0
        self.batches_run = 0
        self._pool: Optional[ThreadPoolExecutor] = None
        self._pending = []              # [(future, jobs, sv table)] of the batches that are on the GPU, oldest first
        self._engines: List[Engine] = []

    # ---- reference-shaped interface
    def initialize(self, key, is_sbnd=False):
        self.temp_key = key
        L = self.validate_sequence_length
        if is_sbnd:
            var1, ref1 = sbnd_templates(key, self.reference_fasta_h, L)
            self._job = SvJob(key, True, [var1, None, ref1, None], False)
        else:
            var1, var2, ref1, ref2, short = canonical_templates(key, self.reference_fasta_h, L)
            self._job = SvJob(key, False, [var1, var2, ref1 if short else None, ref2 if short else None], short)

    def add_long_read_seq(self, treadid, tseq, tmapq1, tmapq2):
        job = self._job
        rid_tokens = treadid.split()
        rid = rid_tokens[0] if rid_tokens else ""
        seq = tseq.strip()
        job.read_mapq[rid] = (_mapq(tmapq1), _mapq(tmapq2))
        if not seq:
            return          # pyssw.read drops records with an empty sequence (pyssw.py:39-40)
        if rid not in job.read_seqs:
            job.read_ids.append(rid)
        else:
            job.bases -= len(job.read_seqs[rid])
        job.read_seqs[rid] = seq
        job.bases += len(seq)

    def _add_encoded_read(self, rid_b: bytes, codes, mapq1_b: bytes, mapq2_b: Optional[bytes]):
        """add_long_read_seq for the byte-level reader: fields as bytes; the sequence, already stripped, as a code array
        or as an AsciiSlice of the block it was read in."""
        job = self._job
        tok = rid_b.decode().split()
        rid = tok[0] if tok else ""
        job.read_mapq[rid] = (None if mapq1_b == b"None" else int(mapq1_b),
                              None if mapq2_b is None or mapq2_b == b"None" else int(mapq2_b))
        n = len(codes)
        if n == 0:
            return
        if rid not in job.read_seqs:
            job.read_ids.append(rid)
        else:
            job.bases -= len(job.read_seqs[rid])
        job.read_seqs[rid] = codes
        job.bases += n

    def count_alignment_and_print(self):
        self._enqueue()

    def count_alignment_and_print_sbnd(self):
        self._enqueue()

    def _enqueue(self):
        if self._job is None:
            return
        self._queue.append(self._job)
        self._queued_bases += self._job.bases
        self._job = None
        if self._queued_bases >= BATCH_BASES:
            self._submit()

    # ---- GPU batch
    def flush(self):
        """Send what is queued to the GPU and write every pending count / info line."""
        self._submit()
        self._drain()

    def _submit(self):
        if not self._queue:
            return
        eng = self._engine or default_engine()
        bb = BatchBuilder(self.score_ratio_thres, self.start_pos_thres, self.end_pos_thres,
                          self.var_ref_margin_thres, self.var_read_min_mapq)
        shared: Dict[int, int] = {}      # id(sequence object) -> index in this batch's pack: the direct gather hands the
        prune = getattr(eng, "prune", True)
        q = self.var_read_min_mapq
        for job in self._queue:          # SAME str object to every SV candidate a read serves, so it is packed once
            bb.add_sv(job.templates, job.is_sbnd, job.short)
            for rid in job.read_ids:
                m1, m2 = job.read_mapq.get(rid, (None, None))
                seq = job.read_seqs[rid]
                # a read whose mapq test fails is only counted (:344-346, :388-389): neither encoded nor uploaded
                if prune and (m1 is None or m1 < q or (not job.is_sbnd and (m2 is None or m2 < q))):
                    bb.add_read_absent(len(seq), m1, m2)
                    continue
                at = shared.get(id(seq)) if isinstance(seq, str) else None
                if at is None:
                    r = bb.add_read(seq, m1, m2)
                    if isinstance(seq, str):
                        shared[id(seq)] = bb.read_seq_index(r)
                else:
                    bb.add_read_shared(at, m1, m2)
        pack, svs, reads = bb.finalize()
        queue, self._queue, self._queued_bases = self._queue, [], 0
        self._launch(pack, svs, reads, queue)

    def _launch(self, pack, svs, reads, metas):
        """The GPU call runs on a worker thread (ctypes releases the GIL), so the caller goes on parsing and encoding the
        next batch meanwhile; results are written strictly in submission order (at the next flush, or at close())."""
        if self._pool is None:
            # Two calls in flight: while the GPU finishes one batch (the end of a launch runs at low occupancy) the next
            # one is uploaded and started. A context serves one call at a time, so the second call goes to a second
            # context on the same GPU; it is created on first use and closed with the counter.
            eng = self._engine or default_engine()
            self._engines = [eng]
            if isinstance(eng, Engine):
                self._engines += _spare_engines(eng, max(1, INFLIGHT) - 1)
            self._pool = ThreadPoolExecutor(max_workers=len(self._engines))
        while len(self._pending) >= len(self._engines):
            self._drain_one()
        eng = self._engines[self.batches_run % len(self._engines)]
        import time
        t_sub = time.perf_counter()
        scoring = self.scoring       # the closure must not hold `self`: a counter whose last reference dies on a worker
                                     # thread would run close() there and wait for that very thread

        def call(eng=eng, k=self.batches_run):
            # `pack` may be a callable that encodes and packs the sequences: that work (C, GIL released) then runs on the
            # worker thread too, beside the caller's parsing
            t0 = time.perf_counter()
            the_pack = pack() if callable(pack) else pack
            t1 = time.perf_counter()
            out = eng.validate(the_pack, svs, reads, scoring)
            if _TRACE:
                print(f"[stage] batch {k}: {len(svs)} SVs {int(the_pack.finalize()[0].size) >> 20} MB queued "
                      f"{1e3 * (t0 - t_sub):.1f} ms, pack {1e3 * (t1 - t0):.1f} ms, GPU call {1e3 * (time.perf_counter() - t1):.1f} ms, "
                      f"done at {1e3 * (time.perf_counter() - _T0):.1f}", flush=True)
            return out
        if _TRACE:
            print(f"[stage] batch {self.batches_run} submitted at {1e3 * (t_sub - _T0):.1f} ms", flush=True)
        fut = self._pool.submit(call)
        self._pending.append((fut, metas, svs))
        self.batches_run += 1

    def _submit_scanned(self, data, src, groups: np.ndarray, sreads: np.ndarray, is_sbnd: bool):
        """A batch straight from the tables of nmsv_seam_scan: per SV candidate only the templates are built in Python
        (reference :207-283); the read table is assembled with numpy, the sequences are encoded and packed by the
        library from the mapping of the seam file, and reads whose mapq test already fails are neither encoded nor
        uploaded (they are only counted, :344-346 / :388-389)."""
        from .my_seq import encode
        if self._queue:
            self._submit()           # keep the output order when the class is also driven by hand
        L = self.validate_sequence_length
        ng, nr = len(groups), len(sreads)
        svs = np.zeros(ng, dtype=_lib.SV_DTYPE)
        svs["tmpl"] = _lib.NMSV_NONE
        svs["tmpl_rc"] = _lib.NMSV_NONE
        svs["read_begin"], svs["read_end"] = groups["read_begin"], groups["read_end"]
        svs["margin"], svs["min_mapq"] = self.var_ref_margin_thres, self.var_read_min_mapq
        tmpl_codes: List[np.ndarray] = []
        metas = []
        for g in range(ng):
            ko, kl = int(groups["key_off"][g]), int(groups["key_len"][g])
            key = data[ko:ko + kl].decode()
            if is_sbnd:
                var1, ref1 = sbnd_templates(key, self.reference_fasta_h, L)
                templates, short = [var1, None, ref1, None], False
            else:
                var1, var2, ref1, ref2, short = canonical_templates(key, self.reference_fasta_h, L)
                templates = [var1, var2, ref1 if short else None, ref2 if short else None]
            seen: Dict[str, Tuple[int, int]] = {}
            row = svs[g]
            for slot, t in enumerate(templates):
                if t is None:
                    continue
                if not t:
                    raise ValueError("empty template")
                if t not in seen:      # identical templates of one SV (var1 == var2 without an insertion) share work
                    rc = _lib.NMSV_NONE
                    if needs_explicit_revcomp(t):
                        rc = len(tmpl_codes); tmpl_codes.append(encode(reverse_complement(t)))
                    seen[t] = (len(tmpl_codes), rc)
                    tmpl_codes.append(encode(t))
                row["tmpl"][slot], row["tmpl_rc"][slot] = seen[t]
            for v in (0, 1):
                if templates[v] is not None:
                    b = integer_bounds(len(templates[v]), self.score_ratio_thres, self.start_pos_thres, self.end_pos_thres)
                    row["score_min"][v], row["qstart_max"][v], row["qend_min"][v] = b
            row["flags"] = (_lib.SV_SBND if is_sbnd else 0) | (_lib.SV_SHORT_DEL_DUP if short else 0)
            metas.append(_ScannedSv(key, is_sbnd, short, data, sreads, int(groups["read_begin"][g])))
        # ---- sequence tables: templates first, then one entry per read
        nt = len(tmpl_codes)
        lens = np.empty(nt + nr, dtype=np.uint32)
        lens[:nt] = [len(c) for c in tmpl_codes]
        lens[nt:] = sreads["seq_len"]
        m1, m2 = sreads["mapq1"], sreads["mapq2"]
        keep = (m1 != _lib.MAPQ_NONE) & (m1 >= self.var_read_min_mapq)
        if not is_sbnd:
            keep &= (m2 != _lib.MAPQ_NONE) & (m2 >= self.var_read_min_mapq)
        if not getattr(eng_prune_probe(self), "prune", True):
            keep[:] = True
        def build_pack():        # runs on the worker thread of this batch (numpy and C: the GIL is released for most of it)
            present = np.ones(nt + nr, dtype=bool)
            present[nt:] = keep
            padded = np.where(present, (lens.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15), np.uint64(0))
            offsets = np.zeros(nt + nr, dtype=np.uint64)
            if nt + nr > 1:
                np.cumsum(padded[:-1], out=offsets[1:])
            total = int(offsets[-1] + padded[-1]) if nt + nr else 0
            codes = np.empty(total, dtype=np.uint8)          # every byte is written below (sequences + their padding)
            for i, c in enumerate(tmpl_codes):
                o = int(offsets[i])
                codes[o:o + len(c)] = c
                codes[o + len(c):o + int(padded[i])] = 4
            kidx = np.flatnonzero(keep)
            if len(kidx):
                so = np.ascontiguousarray(sreads["seq_off"][kidx])
                ln = np.ascontiguousarray(sreads["seq_len"][kidx])
                do = np.ascontiguousarray(offsets[nt + kidx])
                _lib.load().nmsv_encode_pack_padded(src.ctypes.data, so.ctypes.data, ln.ctypes.data, do.ctypes.data, len(kidx),
                                                    codes.ctypes.data)
            offsets[nt:][~keep] = _lib.SEQ_ABSENT
            return SeqPack.from_arrays(codes, offsets, lens)

        reads = np.zeros(nr, dtype=_lib.READ_DTYPE)
        reads["seq"] = nt + np.arange(nr, dtype=np.uint32)
        reads["sv"] = np.repeat(np.arange(ng, dtype=np.uint32), (groups["read_end"] - groups["read_begin"]).astype(np.int64))
        reads["mapq1"], reads["mapq2"] = m1, (_lib.MAPQ_NONE if is_sbnd else m2)
        self._launch(build_pack, svs, reads, metas)

    def _drain(self):
        """Wait for the batches in flight and write their count / info lines, in submission order."""
        while self._pending:
            self._drain_one()

    def _drain_one(self):
        fut, queue, svs = self._pending.pop(0)
        checked, supporting, records = fut.result()       # re-raises an EngineError of the worker
        # plain Python ints in three bulk conversions (per-record numpy scalar access is ~20 us per record)
        n_rec = len(records)
        rec_sv = records["sv"].tolist()
        rec_read = records["read"].tolist()
        rec_aln = np.ascontiguousarray(records["aln"]).view(np.int32).reshape(n_rec, 24).tolist() if n_rec else []
        begin = svs["read_begin"].tolist()
        checked_l, supporting_l = checked.tolist(), supporting.tolist()
        by_sv: Dict[int, list] = {}
        for k in range(n_rec):
            by_sv.setdefault(rec_sv[k], []).append(k)
        none12 = "\t".join(["None"] * 12)
        count_lines, info_lines = [], []

        def six(a, o):        # one entry of the dict of ssw_check (:162): score, qstart, qend, tstart, tend, strand
            return f"{a[o]}\t{a[o + 1]}\t{a[o + 2]}\t{a[o + 3]}\t{a[o + 4]}\t{'+' if a[o + 5] > 0 else '-'}"
        for i, job in enumerate(queue):
            head = job.key.replace(",", "\t")
            count_lines.append(f"{head}\t{checked_l[i]}\t{supporting_l[i]}\n")
            for k in by_sv.get(i, ()):
                a = rec_aln[k]
                rid = job.read_id(rec_read[k] - begin[i])
                if job.is_sbnd:
                    tail = six(a, 0)
                elif job.short:
                    tail = f"{six(a, 0)}\t{six(a, 6)}\t{six(a, 12)}\t{six(a, 18)}"
                else:
                    tail = f"{six(a, 0)}\t{six(a, 6)}\t{none12}"
                info_lines.append(f"{head}\t{rid}\t{tail}\n")
        self.hout_count.write("".join(count_lines))
        self.hout_ainfo.write("".join(info_lines))

    def close(self):
        self.flush()
        if self._pool is not None:
            import threading
            on_worker = threading.current_thread() in getattr(self._pool, "_threads", ())      # (garbage collection there)
            self._pool.shutdown(wait=not on_worker)
            self._pool = None
            self._engines = []          # the extra contexts stay in the per-device pool for the next counter
        for h in (self.hout_count, self.hout_ainfo):
            if not h.closed:
                h.close()
        if hasattr(self.reference_fasta_h, "close"):
            try:
                self.reference_fasta_h.close()
            except Exception:
                pass

    # The reference writes every line inside count_alignment_and_print; here lines are written when a batch comes back
    # from the GPU, so a caller that drives the class by hand must end with close() (or use it as a context manager).
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        unfinished = bool(getattr(self, "_queue", None)) or bool(getattr(self, "_pending", None))
        try:
            self.close()
        except Exception as exc:      # an error of the GPU batch must not vanish with the object
            import warnings
            warnings.warn(f"Alignment_counter for {getattr(self.hout_count, 'name', '?')} was dropped with unwritten "
                          f"results and closing it failed: {exc!r}; the count / info files are incomplete", RuntimeWarning)
        else:
            if unfinished:
                import warnings
                warnings.warn("Alignment_counter was dropped without close(): its last batch was only written at "
                              "garbage collection", RuntimeWarning)


_SPARES: Dict[int, List[Engine]] = {}


def _spare_engines(eng: Engine, n: int) -> List[Engine]:
    """Extra contexts on eng's GPU for calls in flight; created once per process and device (a context allocates its
    rings and snapshot scratch), shared by successive counters (tumor, control, single breakends ...)."""
    pool = _SPARES.setdefault(eng.device, [])
    pool[:] = [e for e in pool if getattr(e, "_h", None)]
    while len(pool) < n:
        pool.append(Engine(eng.device))
    for e in pool[:n]:
        if e.prune != eng.prune:
            e.set_option("prune", 1 if eng.prune else 0)
    return pool[:n]


_TRACE = os.environ.get("NMSV_STAGE_TRACE") is not None      # per-batch timeline on stdout (measurement knob)
import time as _time
_T0 = _time.perf_counter()


def eng_prune_probe(counter) -> Engine:
    return counter._engine or default_engine()


_WS = b" \t\r\n\x0b\x0c"


def _stream_seam_native(seam_file: str, counter: Alignment_counter, is_sbnd: bool) -> None:
    """The group-by of the reference's loop (:414-425 / :436-448) done by the library in one pass over a mapping of the
    seam file (nmsv_seam_scan): Python sees one table of key groups and one of reads per device batch, builds the
    templates of each candidate, and hands the batch to the GPU while the next one is scanned. The line-by-line reader
    (NMSV_SEAM_READER=lines) stays the executable specification; tests hold the readers to identical outputs."""
    import ctypes
    import mmap
    L = _lib.load()
    with open(seam_file, "rb") as hin:
        size = os.fstat(hin.fileno()).st_size
        if size:
            data = mmap.mmap(hin.fileno(), 0, access=mmap.ACCESS_READ)
            src = np.frombuffer(data, dtype=np.uint8)
            # The scan is pure C with the GIL released: it runs one batch ahead on a thread of its own while this thread
            # builds templates and tables for the previous batch.
            import queue
            import threading
            q: "queue.Queue" = queue.Queue(maxsize=2)
            stop = threading.Event()

            def scanner():
                try:
                    gcap, rcap = 4096, 1 << 17
                    groups = np.zeros(gcap, dtype=_lib.SEAM_GROUP_DTYPE)
                    sreads = np.zeros(rcap, dtype=_lib.SEAM_READ_DTYPE)
                    ng, nr = ctypes.c_uint32(0), ctypes.c_uint32(0)
                    nxt, err = ctypes.c_uint64(0), ctypes.c_uint64(0)
                    pos = 0
                    n_batches = 0
                    while pos < size and not stop.is_set():
                        # the first batches are smaller: the GPU starts after a quarter of a batch has been parsed
                        want = max(BATCH_BASES >> max(2 - n_batches, 0), 1)
                        rc = L.nmsv_seam_scan(src.ctypes.data, size, pos, want, BATCH_BASES // 2, 1 if is_sbnd else 0,
                                              groups.ctypes.data, gcap, sreads.ctypes.data, rcap, ctypes.byref(ng),
                                              ctypes.byref(nr), ctypes.byref(nxt), ctypes.byref(err))
                        if rc == _lib.E_CAPACITY:          # one candidate with more reads than the table holds
                            rcap *= 2
                            sreads = np.zeros(rcap, dtype=_lib.SEAM_READ_DTYPE)
                            continue
                        if rc != 0:
                            line_no = data[:err.value].count(b"\n") + 1
                            raise ValueError(f"{seam_file}: line {line_no} is not `key, read id, mapq1, mapq2, sequence` (5 "
                                             "tab-separated fields, integer or None mapping qualities)")
                        if ng.value:
                            q.put((groups[:ng.value].copy(), sreads[:nr.value].copy()))
                            n_batches += 1
                        pos = nxt.value
                    q.put(None)
                except BaseException as exc:      # handed to the consumer
                    q.put(exc)

            th = threading.Thread(target=scanner, daemon=True)
            th.start()
            try:
                while True:
                    item = q.get()
                    if item is None:
                        break
                    if isinstance(item, BaseException):
                        raise item
                    counter._submit_scanned(data, src, item[0], item[1], is_sbnd)
            finally:                    # on an error: let the scanner run out instead of blocking on a full queue
                stop.set()
                while th.is_alive():
                    try:
                        q.get(timeout=0.05)
                    except queue.Empty:
                        pass
                th.join()
    counter.close()          # waits for the batches in flight: their tables keep the mapping alive until then


def _stream_seam_fast(seam_file: str, counter: Alignment_counter, is_sbnd: bool) -> None:
    """The same group-by as _stream_seam, on bytes: the file is memory-mapped, the five fields of a line are located with
    find (memchr), and every read reaches the batch as a slice of the mapping; the library encodes and packs all of them
    in one pass when the batch is finalized (nmsv_encode_pack) -- no per-read str objects, no per-read encoding, no
    intermediate copies. The line-by-line reader stays as the executable specification
    (tests/test_abi_and_host.py holds the two to identical batches)."""
    import mmap

    def close_group():
        if counter.temp_key is not None:
            counter.count_alignment_and_print_sbnd() if is_sbnd else counter.count_alignment_and_print()

    with open(seam_file, "rb") as hin:
        size = os.fstat(hin.fileno()).st_size
        if size:
            data = mmap.mmap(hin.fileno(), 0, access=mmap.ACCESS_READ)
            src = np.frombuffer(data, dtype=np.uint8)          # the reads stay ASCII in the mapping until a batch is packed
            find = data.find
            cur_key = None                                     # bytes of the key column of the open group
            pos = 0
            while pos < size:
                le = find(b"\n", pos)
                if le < 0:
                    le = size                                  # last line without a newline
                t0 = find(b"\t", pos, le)
                t1 = find(b"\t", t0 + 1, le)
                t2 = find(b"\t", t1 + 1, le)
                t3 = find(b"\t", t2 + 1, le)
                if t0 < 0 or t1 < 0 or t2 < 0 or t3 < 0 or find(b"\t", t3 + 1, le) >= 0:
                    raise ValueError(f"{seam_file}: a line does not have 5 tab-separated fields")
                kb = data[pos:t0]
                if kb != cur_key:
                    close_group()
                    counter.initialize(kb.decode(), is_sbnd=is_sbnd)
                    cur_key = kb
                a, b = t3 + 1, le
                while b > a and data[b - 1] in _WS:            # str.strip() of the line-by-line reader (e.g. CR LF files)
                    b -= 1
                while a < b and data[a] in _WS:
                    a += 1
                counter._add_encoded_read(data[t0 + 1:t1], AsciiSlice(src, a, b - a), data[t1 + 1:t2],
                                          None if is_sbnd else data[t2 + 1:t3])
                pos = le + 1
    close_group()
    counter.close()          # waits for the batches in flight: their slices keep the mapping alive until then


def _stream_seam(seam_file: str, counter: Alignment_counter, is_sbnd: bool) -> None:
    """Group-by over the key-sorted seam TSV (reference :414-425 and :436-448)."""
    with open(seam_file) as hin:
        for line in hin:
            tkey, treadid, tmapq1, tmapq2, tseq = line.rstrip("\n").split("\t")
            if is_sbnd:
                tmapq2 = "None"
            if counter.temp_key != tkey:
                if counter.temp_key is not None:
                    counter.count_alignment_and_print_sbnd() if is_sbnd else counter.count_alignment_and_print()
                counter.initialize(tkey, is_sbnd=is_sbnd)
            counter.add_long_read_seq(treadid, tseq, tmapq1, tmapq2)
        if counter.temp_key is not None:
            counter.count_alignment_and_print_sbnd() if is_sbnd else counter.count_alignment_and_print()
    counter.close()


def count_sread_from_seam(seam_file, output_count_file, output_alignment_info_file, reference_fasta, is_sbnd=False,
                          var_read_min_mapq=0, score_ratio_thres=1.2, debug=False, engine=None, **kwargs):
    """The part of count_sread_by_alignment after read gathering: seam TSV -> sread_count / sread_info files."""
    counter = Alignment_counter(output_count_file, output_alignment_info_file, reference_fasta, var_read_min_mapq,
                                score_ratio_thres, False, debug, engine=engine, **kwargs)
    # NMSV_SEAM_READER=lines selects the line-by-line reader that mirrors the reference's loop, =bytes the Python
    # byte-level reader of round 1; the default is the library's one-pass scan
    reader = {"lines": _stream_seam, "bytes": _stream_seam_fast}.get(os.environ.get("NMSV_SEAM_READER", "native"),
                                                                     _stream_seam_native)
    reader(seam_file, counter, is_sbnd)
    return counter


def count_sread_by_alignment(sv_file, alignment_file, output_count_file, output_alignment_info_file, reference_fasta,
                             sbnd_file=None, output_count_file_sbnd=None, output_alignment_info_file_sbnd=None,
                             check_read_max_num=500, var_read_min_mapq=0, score_ratio_thres=1.2, use_ssw_lib=False,
                             sort_option=None, debug=False, engine: Optional[Engine] = None):
    """Same signature as the reference (:401-403) plus an optional engine handle. By default the gathered reads stay in
    memory (gather_reads_direct: no seam TSV of whole reads per key, no GNU sort of it, no re-parse; a read that serves
    several candidates is packed once; reads whose mapq test fails are not packed at all). With debug=True -- where the
    reference leaves its `.tmp.local_read_for_realignment` file behind -- or NMSV_DIRECT_GATHER=0 the reference's own
    route through the sorted seam file is taken; both routes hand the GPU the same work and write the same files."""
    if os.environ.get("NMSV_DIRECT_GATHER", "0" if debug else "1") == "1":
        for stale in (output_count_file, output_count_file_sbnd):      # the reference rewrites and removes its seam files
            if stale is not None and os.path.exists(stale + ".tmp.local_read_for_realignment"):
                os.remove(stale + ".tmp.local_read_for_realignment")
        groups, groups_sbnd = gather_reads_direct(sv_file, alignment_file, reference_fasta, sbnd_file=sbnd_file,
                                                  sort_option=sort_option)
        _count_from_groups(groups, Alignment_counter(output_count_file, output_alignment_info_file, reference_fasta,
                                                     var_read_min_mapq, score_ratio_thres, False, debug, engine=engine), False)
        if sbnd_file is not None:
            _count_from_groups(groups_sbnd, Alignment_counter(output_count_file_sbnd, output_alignment_info_file_sbnd,
                                                              reference_fasta, var_read_min_mapq, score_ratio_thres, False,
                                                              debug, engine=engine), True)
        return
    seam = output_count_file + ".tmp.local_read_for_realignment"
    seam_sbnd = (output_count_file_sbnd + ".tmp.local_read_for_realignment") if output_count_file_sbnd is not None else None
    gather_local_read_for_realignment(sv_file, alignment_file, seam, reference_fasta, sbnd_file=sbnd_file,
                                      output_file_sbnd=seam_sbnd, check_read_max_num=check_read_max_num,
                                      sort_option=sort_option)
    count_sread_from_seam(seam, output_count_file, output_alignment_info_file, reference_fasta, False,
                          var_read_min_mapq, score_ratio_thres, debug, engine)
    if not debug:
        os.remove(seam)
    if sbnd_file is None:
        return
    count_sread_from_seam(seam_sbnd, output_count_file_sbnd, output_alignment_info_file_sbnd, reference_fasta, True,
                          var_read_min_mapq, score_ratio_thres, debug, engine)
    if not debug:
        os.remove(seam_sbnd)
