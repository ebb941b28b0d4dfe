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


def _problem(contig, fasta_file_ins, chr1, start1, end1, dir1, chr2, start2, end2, dir2, mode, rd_margin=20, i_margin=500):
    """((first, region1, region2) for sw_jump, context for _finish)"""
    start1 = max(1, start1)
    start2 = max(1, start2)
    if mode != "i":
        ref_len1 = fasta_file_ins.get_reference_length(chr1)
        ref_len2 = fasta_file_ins.get_reference_length(chr2)
        bstart1, bend1 = max(int(start1) - rd_margin, 1), int(end1) + rd_margin
        bstart2, bend2 = max(int(start2) - rd_margin, 1), int(end2) + rd_margin
        if ref_len1 < bend1:
            bend1 = ref_len1
        if ref_len2 < bend2:
            bend2 = ref_len2
        region1_seq = fasta_file_ins.fetch(chr1, bstart1 - 1, bend1).upper()
        region2_seq = fasta_file_ins.fetch(chr2, bstart2 - 1, bend2).upper()
        if dir1 == '-':
            region1_seq = reverse_complement(region1_seq)
        if dir2 == '+':
            region2_seq = reverse_complement(region2_seq)
        return (contig, region1_seq, region2_seq), (mode, contig, dir1, dir2, bstart1, bend1, bstart2, bend2)
    contig_start = contig[:min(i_margin, len(contig))]
    contig_end = contig[-min(i_margin, len(contig)):]
    region_seq = fasta_file_ins.fetch(chr1, max(0, start1 - 100), end2 + 100).upper()
    return (region_seq, contig_start, contig_end), (mode, contig, len(contig_end), max(0, start1 - 100))


def _finish(sret, ctx, h_log):
    if sret is None:
        return None
    if ctx[0] != "i":
        _, contig, dir1, dir2, bstart1, bend1, bstart2, bend2 = ctx
        score, contig_align, region1_align, region2_align, contig_seq, region_seq = sret
        bp_pos1 = bstart1 + region1_align[1] - 1 if dir1 == '+' else bend1 - region1_align[1] + 1
        bp_pos2 = bstart2 + region2_align[0] - 1 if dir2 == '-' else bend2 - region2_align[0] + 1
        if contig_align[2] - contig_align[1] == 1:
            inseq = '---'
        elif contig_align[2] - contig_align[1] > 1:
            inseq = contig[(contig_align[1]):(contig_align[2] - 1)]
            if dir1 == '-':
                inseq = reverse_complement(inseq)
        else:
            logger.warning("Alignment inconsistent!!")
        if h_log is not None:
            print(score, contig_align, region1_align, region2_align, file=h_log)
            print(contig_seq, file=h_log)
            print(region_seq, file=h_log)
        return (bp_pos1, bp_pos2, inseq)      # unassigned after the warning above: fails like the reference (:58)
    _, contig, len_contig_end, origin = ctx
    score, region_align, contig_start_align, contig_end_align, region_seq, contig_seq = sret
    bp_pos1 = origin + region_align[1]
    bp_pos2 = origin + region_align[2]
    inseq_start = contig_start_align[1]
    inseq_end = len(contig) - (len_contig_end - contig_end_align[0] + 1)
    inseq = contig[inseq_start:(inseq_end + 1)]
    if h_log is not None:
        print(score, region_align, contig_start_align, contig_end_align, file=h_log)
        print(region_seq, file=h_log)
        print(contig_seq, file=h_log)
    return (bp_pos1, bp_pos2, inseq)


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
