"""Host logic of the two call-site drop-ins (locate_bp_sbnd, generate_paf_file) without a GPU: the alignment batch is answered by
the CPU oracle (tests only), everything around it -- windows on the reference, clipping, strand handling, file parsing, position
arithmetic, output / log / PAF formats -- must reproduce the files written by the reference's own functions
(tests/golden/callsites_reference.json, generator make_golden_callsites.py). The GPU twin is tests/test_gpu_callsites.py."""
import json
import os

import pytest

from nanomonsv_b200 import generate_consensus as GC, locate_bp_sbnd as LB, ssw as SSW
from oracle import oracle as O

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "callsites_reference.json")))


class _Fasta:
    def __init__(self, contigs):
        self.contigs = contigs

    def fetch(self, chrom, start, end):
        return self.contigs[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(self.contigs[chrom])


@pytest.fixture
def oracle_backed(monkeypatch):
    def ssw_batch(pairs, open, extend, matrix, engine=None):
        return [O.ssw(s1, s2, open, extend, matrix.match, matrix.mismatch) for s1, s2 in pairs]
    monkeypatch.setattr(SSW, "ssw_batch", ssw_batch)
    monkeypatch.setattr(SSW, "ssw", lambda s1, s2, o, e, m, engine=None: ssw_batch([(s1, s2)], o, e, m)[0])


def test_locate_bp_sbnd_host_logic(oracle_backed, tmp_path):
    d = FIX["sbnd"]
    cfile, out = tmp_path / "consensus.txt", tmp_path / "out.txt"
    cfile.write_text("\n".join(d["consensus_lines"]) + "\n")
    LB.locate_bp_sbnd(str(cfile), str(out), _Fasta(d["contigs"]), True)
    assert out.read_text().splitlines() == d["output"]
    assert open(str(out) + ".locate_bp.sbnd.log").read().splitlines() == d["log"]
    LB.locate_bp_sbnd(str(cfile), str(out), _Fasta(d["contigs"]), False)
    assert not os.path.exists(str(out) + ".locate_bp.sbnd.log")
    key, cons = d["consensus_lines"][3].split("\t")          # a window clipped at the contig end, '-' strand
    tchr, tstart, tend, tdir, _, tcid = key.split(",")
    bp, after = LB.get_refined_bp_sbnd(cons, _Fasta(d["contigs"]), tchr, int(tstart), int(tend), tdir, None)
    assert f"{tchr}\t{bp}\t{tdir}\t{after}\tb_{tcid}" == d["output"][3]


def test_generate_paf_file_host_logic(oracle_backed, tmp_path):
    d = FIX["paf"]
    qf, tf, of = tmp_path / "q.fa", tmp_path / "t.fa", tmp_path / "o.paf"
    qf.write_text(d["query_fasta"]); tf.write_text(d["target_fasta"])
    errors = []
    assert GC.generate_paf_file(str(qf), str(tf), str(of), errors) == d["returned"]
    assert of.read_text().splitlines() == d["paf"] and errors == []
    # a pair without a result (parasail: NULL on 16-bit saturation) gets no line and is reported like the reference does (:123-125)
    import nanomonsv_b200.ssw as S
    real = S.ssw_batch
    S.ssw_batch = lambda pairs, o, e, m, engine=None: [None if k == 1 else r for k, r in enumerate(real(pairs, o, e, m))]
    try:
        errors = []
        n = GC.generate_paf_file(str(qf), str(tf), str(of), errors)
    finally:
        S.ssw_batch = real
    assert n == d["returned"] - 1 and errors == [("read1", "cons_1")]
    assert of.read_text().splitlines() == d["paf"][:1] + d["paf"][2:]


class _Sv:
    def __init__(self, c1, p1, d1, c2, p2, d2, inseq, sid, flt):
        self.chr1, self.pos1, self.dir1, self.chr2, self.pos2, self.dir2, self.inseq, self.sv_id = c1, p1, d1, c2, p2, d2, inseq, sid
        self.filter = list(flt)


def test_insertion_match_filter_host_logic(oracle_backed):
    """Sv_filterer.filter_sv_insertion_match in the pair order of apply_filters (post_proc.py:84-134, :226-232): the filters
    the reference's own method left on its own Sv objects."""
    from nanomonsv_b200 import sv_filter as F
    d = FIX["filter"]
    svs = [_Sv(*row) for row in d["svs"]]
    n = F.apply_insertion_match_filter(svs, _Fasta(d["contigs"]))
    assert [s.filter for s in svs] == d["filters_after"]
    assert n > 0 and n % 2 == 0
    assert any(f == ["Duplicate_with_insertion"] for f in d["filters_after"]) and any(f == [] for f in d["filters_after"])
    # nothing to examine: no alignment call at all
    assert F.apply_insertion_match_filter([_Sv("chrA", 10, "+", "chrA", 500, "-", "", "x", [])], _Fasta(d["contigs"])) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/nanomonsv"), reason="the reference checkout is not on this box")
def test_the_fixture_is_what_its_generator_writes():
    """tests/golden/make_golden_callsites.py re-run here (reference code from /root/reference, oracle-backed parasail stand-in)."""
    import sys
    golden = os.path.join(os.path.dirname(__file__), "golden")
    saved = {k: sys.modules[k] for k in list(sys.modules) if k in ("pysam", "parasail") or k == "nanomonsv" or k.startswith("nanomonsv.")}
    sys.path.insert(0, golden)
    try:
        import make_golden_callsites as M
        got = M.build()
        for part in ("sbnd", "paf", "filter"):
            assert got[part] == FIX[part], part
    finally:
        sys.path.remove(golden)
        for k in [k for k in sys.modules if k in ("pysam", "parasail") or k == "nanomonsv" or k.startswith("nanomonsv.")]:
            del sys.modules[k]
        sys.modules.update(saved)
