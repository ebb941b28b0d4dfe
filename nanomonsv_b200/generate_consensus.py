"""GPU-backed drop-in for the alignment step of nanomonsv/generate_consensus.py (reference v0.8.0):
`Consensus_generator.generate_paf_file` (:100-127) aligns every supporting read of a candidate against the first consensus
with `parasail.ssw(qseq, tseq, 3, 1, matrix_create("ACGT", 2, -2))` -- s1 is the WHOLE read (10-50 kb), s2 the consensus --
and writes one PAF line per read for racon. Here all reads of a candidate are one GPU batch (nmsv_ssw_batch through the
parasail-shaped facade). SURVEY.md section 8f row 3.

Kept as in the reference: the target is the LAST record of the target file, its id cut at the first blank (:105-110); a query
id is the whole header line without '>' (:115); one sequence line per record; a pair for which parasail returns no result
(16-bit score saturation, never a matter of length) is appended to `parasail_error` and gets no line (:120-125); the
columns of the line (:121-122); the return value is the number of lines written."""
from __future__ import annotations

from typing import List, Optional, Tuple

from . import ssw as _ssw
from .engine import Engine


def generate_paf_file(query_fasta, target_fasta, output_file, parasail_error: Optional[list] = None,
                      engine: Optional[Engine] = None) -> int:
    """Module-level form of the method: bind it with
        def generate_paf_file(self, q, t, o): return nmsv_gc.generate_paf_file(q, t, o, self.parasail_error)"""
    tid = tseq = None
    with open(target_fasta, 'r') as hin:
        for line in hin:
            if line.startswith('>'):
                tid = line.rstrip('\n').split(' ')[0].lstrip('>')
            else:
                tseq = line.rstrip('\n')
    queries: List[Tuple[str, str]] = []
    qid = None
    with open(query_fasta, 'r') as hin:
        for line in hin:
            if line.startswith('>'):
                qid = line.rstrip('\n').lstrip('>')
            else:
                queries.append((qid, line.rstrip('\n')))
    results = _ssw.ssw_batch([(qseq, tseq) for _, qseq in queries], 3, 1, _ssw.matrix_create("ACGT", 2, -2), engine=engine)
    paf_rec_count = 0
    with open(output_file, 'w') as hout:
        for (qid, qseq), res in zip(queries, results):
            if res is not None:
                print(f"{qid}\t{len(qseq)}\t{res.read_begin1}\t{res.read_end1}\t+\t" +
                      f"{tid}\t{len(tseq)}\t{res.ref_begin1}\t{res.ref_end1}\t*\t*\t60", file=hout)
                paf_rec_count = paf_rec_count + 1
            elif parasail_error is not None:
                parasail_error.append((qid, tid))
    return paf_rec_count
