"""GPU-backed drop-in for nanomonsv/locate_bp.py (reference v0.8.0): single-base breakpoints and inserted sequence of every
consensus contig by `smith_waterman.sw_jump` (locate_bp.py:37,67). The reference runs its
pure-Python sw_jump once per consensus line; `locate_bp` here sends all lines of a consensus file to the GPU as one
`sw_jump_batch` (csrc/nmsv_jump.cuh) and writes the same output and log files.

Kept as in the reference: start positions clamped to 1 (:15-16); modes d / r / t: +-20 bp regions, clipped to the contig
lengths, reverse-complemented for dir1 '-' / dir2 '+', the contig as first sequence (:20-39), positions and inserted sequence
from the returned tuples (:44-52; an alignment whose two contig fragments overlap leaves `inseq` unassigned there and fails
at the return -- same here); mode i: the reference window is the first sequence, the first / last 500 contig bases the
regions (:61-78); the log lines (:54-56, :80-82, :100-108)."""
from __future__ import annotations

import logging
import os
from typing import List, Optional, Tuple

from . import smith_waterman
from .engine import Engine
from .my_seq import reverse_complement

logger = logging.getLogger(__name__)


def _window(fasta, chrom, lo, hi, margin):
    """1-based inclusive [lo - margin, hi + margin] clipped to the contig (:23-29): (first, last, upper-case bases)."""
    first = max(int(lo) - margin, 1)
    last = min(int(hi) + margin, fasta.get_reference_length(chrom))
    return first, last, fasta.fetch(chrom, first - 1, last).upper()


def _problem(contig, fasta, chr1, start1, end1, dir1, chr2, start2, end2, dir2, mode, rd_margin=20, i_margin=500):
    """(the three sequences for sw_jump in the order the reference passes them, what _finish needs besides the result)"""
    start1, start2 = max(1, start1), max(1, start2)                              # :15-16
    if mode == "i":
        # insertion: the reference window plays the contig role, the two ends of the contig are the regions (:61-67)
        n_end = min(i_margin, len(contig))
        origin = max(0, start1 - 100)
        window = fasta.fetch(chr1, origin, end2 + 100).upper()
        return (window, contig[:n_end], contig[-n_end:]), ("i", contig, n_end, origin)
    first1, last1, reg1 = _window(fasta, chr1, start1, end1, rd_margin)
    first2, last2, reg2 = _window(fasta, chr2, start2, end2, rd_margin)
    # both regions are presented in the direction the contig runs through them (:34-35)
    if dir1 == '-':
        reg1 = reverse_complement(reg1)
    if dir2 == '+':
        reg2 = reverse_complement(reg2)
    return (contig, reg1, reg2), (mode, contig, dir1, dir2, first1, last1, first2, last2)


def _finish(sret, ctx, h_log):
    """(bp_pos1, bp_pos2, inseq) from a sw_jump result (:41-58, :69-84), None when sw_jump found nothing; writes the three log
    lines of the reference."""
    if sret is None:
        return None
    score, first_aln, r1_aln, r2_aln, row_a, row_b = sret
    if h_log is not None:
        print(score, first_aln, r1_aln, r2_aln, file=h_log)
        print(row_a, file=h_log)
        print(row_b, file=h_log)
    if ctx[0] == "i":
        _, contig, n_end, origin = ctx
        # first_aln = (start, end of the first fragment, start, end of the second) on the reference window
        ins_from = r1_aln[1]
        ins_to = len(contig) - (n_end - r2_aln[0] + 1)
        return origin + first_aln[1], origin + first_aln[2], contig[ins_from:ins_to + 1]
    _, contig, dir1, dir2, first1, last1, first2, last2 = ctx
    # the last aligned base of region1 and the first aligned base of region2, back in forward coordinates
    pos1 = first1 + r1_aln[1] - 1 if dir1 == '+' else last1 - r1_aln[1] + 1
    pos2 = first2 + r2_aln[0] - 1 if dir2 == '-' else last2 - r2_aln[0] + 1
    between = first_aln[2] - first_aln[1]                   # contig bases between the two fragments, plus one
    if between == 1:
        inseq = '---'
    elif between > 1:
        inseq = contig[first_aln[1]:first_aln[2] - 1]
        if dir1 == '-':
            inseq = reverse_complement(inseq)
    else:
        # overlapping fragments: the reference only warns here and then fails at its return statement, inseq never assigned
        logger.warning("Alignment inconsistent!!")
        raise UnboundLocalError("cannot access local variable 'inseq' where it is not associated with a value")
    return pos1, pos2, inseq


def get_refined_bp(contig, fasta_file_ins, chr1, start1, end1, dir1, chr2, start2, end2, dir2, mode, h_log, sw_jump_params,
                   rd_margin=20, i_margin=500, engine: Optional[Engine] = None):
    """Same arguments and return value as the reference (:13-84); a batch of one."""
    trio, ctx = _problem(contig, fasta_file_ins, chr1, start1, end1, dir1, chr2, start2, end2, dir2, mode, rd_margin, i_margin)
    sret = smith_waterman.sw_jump(*trio, match_score=sw_jump_params[0], mismatch_penalty=sw_jump_params[1],
                                  gap_cost=sw_jump_params[2], jump_cost=sw_jump_params[3], engine=engine)
    return _finish(sret, ctx, h_log)


def locate_bp(consensus_file, output_file, reference_fasta, sw_jump_params, debug, engine: Optional[Engine] = None):
    """Same signature, output file and log file as the reference (:88-109). `reference_fasta` is a path (opened with pysam, as
    the reference does) or any object with fetch(chrom, start0, end0) and get_reference_length(chrom)."""
    if isinstance(reference_fasta, str):
        import pysam  # lazy: only this host-side fetch needs it
        fasta_file_ins = pysam.FastaFile(reference_fasta)
    else:
        fasta_file_ins = reference_fasta
    items: List[Tuple] = []
    with open(consensus_file, 'r') as hin:
        for line in hin:
            F = line.rstrip('\n').split('\t')
            temp_key, tconsensus = F[0], F[1]
            chr1, start1, end1, dir1, chr2, start2, end2, dir2, mode, cid = temp_key.split(',')
            trio, ctx = _problem(tconsensus, fasta_file_ins, chr1, int(start1), int(end1), dir1, chr2, int(start2), int(end2), dir2, mode)
            items.append((temp_key, tconsensus, chr1, dir1, chr2, dir2, mode, cid, trio, ctx))
    srets = smith_waterman.sw_jump_batch([it[8] for it in items], sw_jump_params[0], sw_jump_params[1], sw_jump_params[2],
                                         sw_jump_params[3], engine=engine) if items else []
    with open(output_file, 'w') as hout, open(output_file + ".locate_bp.log", 'w') as hout_log:
        for (temp_key, tconsensus, chr1, dir1, chr2, dir2, mode, cid, _, ctx), sret in zip(items, srets):
            print(temp_key + '\n' + tconsensus, file=hout_log)
            bret = _finish(sret, ctx, hout_log)
            if bret is not None:
                bp_pos1, bp_pos2, inseq = bret
                print(bp_pos1, bp_pos2, inseq, file=hout_log)
                print('', file=hout_log)
                print(f"{chr1}\t{bp_pos1}\t{dir1}\t{chr2}\t{bp_pos2}\t{dir2}\t{inseq}\t{mode}_{cid}", file=hout)
    if not debug:
        os.remove(output_file + ".locate_bp.log")
