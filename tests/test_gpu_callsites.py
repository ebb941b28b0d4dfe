"""two other parasail.ssw call sites of the reference as drop-in functions on the GPU, held to outputs
that the reference's own code produced (tests/golden/callsites_reference.json, generator make_golden_callsites.py:
nanomonsv/locate_bp_sbnd.py and nanomonsv/generate_consensus.py imported unmodified, parasail numbers from the oracle)."""
import json
import os

import pytest

from nanomonsv_b200 import generate_consensus as GC, locate_bp_sbnd as LB

pytestmark = pytest.mark.gpu
FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "callsites_reference.json")))


class _Fasta:
    def __init__(self, contigs):
        self.contigs = contigs

    def fetch(self, chrom, start, end):
        return self.contigs[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(self.contigs[chrom])


def test_locate_bp_sbnd_equals_the_reference_run(engine, tmp_path):
    d = FIX["sbnd"]
    cfile, out = tmp_path / "consensus.txt", tmp_path / "out.txt"
    cfile.write_text("\n".join(d["consensus_lines"]) + "\n")
    fa = _Fasta(d["contigs"])
    LB.locate_bp_sbnd(str(cfile), str(out), fa, True, engine=engine)
    assert out.read_text().splitlines() == d["output"]
    assert open(str(out) + ".locate_bp.sbnd.log").read().splitlines() == d["log"]
    LB.locate_bp_sbnd(str(cfile), str(out), fa, False, engine=engine)          # debug off: the log is removed (:66)
    assert not os.path.exists(str(out) + ".locate_bp.sbnd.log")
    # the per-contig helper (:14-40) gives what the batched stage function wrote
    for line, exp in zip(d["consensus_lines"], d["output"]):
        key, cons = line.split("\t")
        tchr, tstart, tend, tdir, _, tcid = key.split(",")
        bp, after = LB.get_refined_bp_sbnd(cons, fa, tchr, int(tstart), int(tend), tdir, None, engine=engine)
        assert f"{tchr}\t{bp}\t{tdir}\t{after}\tb_{tcid}" == exp


def test_generate_paf_file_equals_the_reference_run(engine, tmp_path):
    d = FIX["paf"]
    qf, tf, of = tmp_path / "q.fa", tmp_path / "t.fa", tmp_path / "o.paf"
    qf.write_text(d["query_fasta"]); tf.write_text(d["target_fasta"])
    errors = []
    n = GC.generate_paf_file(str(qf), str(tf), str(of), errors, engine=engine)
    assert of.read_text().splitlines() == d["paf"]
    assert n == d["returned"] and [list(e) for e in errors] == d["parasail_error"]
    assert any(int(l.split("\t")[1]) > 16383 for l in d["paf"])      # whole reads beyond the old length limit are aligned


class _Sv:
    def __init__(self, c1, p1, d1, c2, p2, d2, inseq, sid, flt):
        self.chr1, self.pos1, self.dir1, self.chr2, self.pos2, self.dir2, self.inseq, self.sv_id = c1, p1, d1, c2, p2, d2, inseq, sid
        self.filter = list(flt)


def test_insertion_match_filter_equals_the_reference_run(engine):
    """post_proc.py:84-134 in the pair order of :226-232, all examined pairs as one GPU batch."""
    from nanomonsv_b200 import sv_filter as F
    d = FIX["filter"]
    svs = [_Sv(*row) for row in d["svs"]]
    n = F.apply_insertion_match_filter(svs, _Fasta(d["contigs"]), engine=engine)
    assert [s.filter for s in svs] == d["filters_after"] and n >= 10
