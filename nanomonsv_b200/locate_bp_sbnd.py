"""GPU-backed drop-in for nanomonsv/locate_bp_sbnd.py (reference v0.8.0): breakpoint localisation of single-breakend
consensus contigs. One `parasail.ssw(qseq, tconsensus, 3, 1, matrix_create("ACGT", 1, -2))` per consensus line in the reference
(locate_bp_sbnd.py:28-29); here all lines of a consensus file go to the GPU as one batch
(nmsv_ssw_batch through the parasail-shaped facade, ssw.py).

Kept as in the reference: the window on the reference genome (:19-26: `margin` bases on the far side, clipped coordinates,
reverse complement for '-'), the whole consensus as the second sequence (the 1000-base prefix computed at :16 is never
used), `None` where parasail returns no result (:30-32), the position arithmetic (:35-38), the output line and the log
file (:49-66, the log is removed unless debug)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

from . import ssw as _ssw
from .engine import Engine
from .my_seq import reverse_complement


def _window(fasta_file_h, tchr: str, tstart: int, tend: int, tdir: str, margin: int) -> Tuple[str, int, int]:
    """The reference window of one consensus contig (:18-26) and the clipped (tstart, tend)."""
    ref_len = fasta_file_h.get_reference_length(tchr)
    if tstart < 1:
        tstart = 1
    if ref_len < tend:
        tend = ref_len
    if tdir == '+':
        qseq = fasta_file_h.fetch(tchr, max(int(tstart) - margin, 0), int(tend)).upper()
    else:
        qseq = fasta_file_h.fetch(tchr, max(int(tstart) - 1, 0), int(tend) + margin).upper()
        qseq = reverse_complement(qseq)
    return qseq, tstart, tend


def _refine(res, qseq: str, tconsensus: str, tstart: int, tend: int, tdir: str):
    if res is None:
        return None
    if tdir == '+':
        bp_pos_reference = tend - (len(qseq) - res.read_end1 - 1)
    else:
        bp_pos_reference = tstart + (len(qseq) - res.read_end1 - 1)
    return bp_pos_reference, tconsensus[(res.ref_end1 + 1):]


def get_refined_bp_sbnd(tconsensus, fasta_file_h, tchr, tstart, tend, tdir, hout_log=None, margin=200,
                        engine: Optional[Engine] = None):
    """Same arguments and return value as the reference (:14-40); a batch of one."""
    qseq, tstart, tend = _window(fasta_file_h, tchr, tstart, tend, tdir, margin)
    res = _ssw.ssw(qseq, tconsensus, 3, 1, _ssw.matrix_create("ACGT", 1, -2), engine=engine)
    return _refine(res, qseq, tconsensus, tstart, tend, tdir)


def locate_bp_sbnd(consensus_file, output_file, reference_fasta, debug, engine: Optional[Engine] = None):
    """Same signature, output file and log file as the reference (:43-66). `reference_fasta` is a path (opened with pysam,
    like the reference does) or any object with fetch(chrom, start0, end0) and get_reference_length(chrom)."""
    if isinstance(reference_fasta, str):
        import pysam  # lazy: only this host-side fetch needs it
        fasta_file_ins = pysam.FastaFile(reference_fasta)
    else:
        fasta_file_ins = reference_fasta
    items: List[tuple] = []
    with open(consensus_file, 'r') as hin:
        for line in hin:
            F = line.rstrip('\n').split('\t')
            temp_key, tconsensus = F[0], F[1]
            tchr, tstart, tend, tdir, _, tcid = temp_key.split(',')
            qseq, tstart, tend = _window(fasta_file_ins, tchr, int(tstart), int(tend), tdir, 200)
            items.append((temp_key, tconsensus, tchr, tstart, tend, tdir, tcid, qseq))
    results = _ssw.ssw_batch([(it[7], it[1]) for it in items], 3, 1, _ssw.matrix_create("ACGT", 1, -2), engine=engine)
    with open(output_file, 'w') as hout, open(output_file + ".locate_bp.sbnd.log", 'w') as hout_log:
        for (temp_key, tconsensus, tchr, tstart, tend, tdir, tcid, qseq), res in zip(items, results):
            print(temp_key + '\n' + tconsensus, file=hout_log)
            bret = _refine(res, qseq, tconsensus, tstart, tend, tdir)
            if bret is not None:
                bp_pos, tconsensus_after = bret
                print(bp_pos, tconsensus_after, file=hout_log)
                print('', file=hout_log)
                print(f"{tchr}\t{bp_pos}\t{tdir}\t{tconsensus_after}\tb_{tcid}", file=hout)
    if not debug:
        os.remove(output_file + ".locate_bp.sbnd.log")
