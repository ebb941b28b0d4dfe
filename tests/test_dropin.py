"""nanomonsv_b200.dropin.install() against the reference package itself (build container only: needs /root/reference): the
stage functions the reference's run.py calls are replaced where run.py looks them up, the reference's own apply_filters loop
drives the batched insertion-match verdicts to the filters its own method produced, and `parasail` resolves to the facade."""
import json
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nanomonsv")), reason="the reference checkout is not on this box")
FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "callsites_reference.json")))


class _Fasta:
    def __init__(self, contigs):
        self.contigs = contigs

    def fetch(self, chrom, start, end):
        return self.contigs[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(self.contigs[chrom])

    def close(self):
        pass


@pytest.fixture
def reference_package(monkeypatch):
    saved = {k: v for k, v in sys.modules.items() if k == "nanomonsv" or k.startswith("nanomonsv.") or k in ("pysam", "parasail", "h5py")}
    for k in saved:
        del sys.modules[k]
    pysam = types.ModuleType("pysam")          # absent in this image; nothing here opens a BAM
    pysam.AlignmentFile = pysam.FastaFile = pysam.TabixFile = object
    sys.modules["pysam"] = pysam
    sys.modules["h5py"] = types.ModuleType("h5py")
    monkeypatch.syspath_prepend(REF)
    yield
    for k in [k for k in sys.modules if k == "nanomonsv" or k.startswith("nanomonsv.") or k in ("pysam", "parasail", "h5py")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_install_replaces_what_run_py_calls(reference_package, monkeypatch):
    from nanomonsv_b200 import dropin, ssw as SSW, count_sread_by_alignment as C, locate_bp as LB, locate_bp_sbnd as LBS
    from oracle import oracle as O
    done = dropin.install()
    import nanomonsv.run as run
    import nanomonsv.post_proc as pp
    import nanomonsv.generate_consensus as gc
    import parasail
    assert parasail is SSW                                   # the reference's `import parasail` resolved to the facade
    assert run.count_sread_by_alignment is C.count_sread_by_alignment
    assert run.locate_bp is LB.locate_bp and run.locate_bp_sbnd is LBS.locate_bp_sbnd
    assert set(done) >= {"count_sread_by_alignment", "locate_bp", "locate_bp_sbnd", "generate_paf_file", "filter_sv_insertion_match", "sw_jump"}
    assert gc.Consensus_generator.generate_paf_file is done["generate_paf_file"]

    # the reference's own Sv objects and its own apply_filters stage loop, with the patched method (alignments: oracle, CPU)
    calls = []

    def ssw_batch(pairs, open, extend, matrix, engine=None):
        calls.append(len(pairs))
        return [O.ssw(s1, s2, open, extend, matrix.match, matrix.mismatch) for s1, s2 in pairs]
    monkeypatch.setattr(SSW, "ssw_batch", ssw_batch)
    d = FIX["filter"]
    f = object.__new__(pp.Sv_filterer)
    f.reference_h, f.min_indel_size, f.bp_dist_margin, f.validate_seg_len = _Fasta(d["contigs"]), 50, 30, 100
    f.min_tumor_VAF, f.simple_repeat_tb, f.is_control, f.sv_list = 0.05, None, False, []
    f.hout = open(os.devnull, "w")
    for c1, p1, d1, c2, p2, d2, inseq, sid, flt in d["svs"]:
        f.add_sv(c1, p1, d1, c2, p2, d2, inseq, sid, 30, 10, None, None)
        f.sv_list[-1].filter = list(flt)
    import itertools
    for a, c in itertools.combinations(f.sv_list, 2):
        if len(a.filter) > 0 or len(c.filter) > 0:
            continue
        if len(a.inseq) < f.min_indel_size and len(c.inseq) >= f.min_indel_size:
            f.filter_sv_insertion_match(a, c)
        elif len(c.inseq) < f.min_indel_size and len(a.inseq) >= f.min_indel_size:
            f.filter_sv_insertion_match(c, a)
    assert [s.filter for s in f.sv_list] == d["filters_after"]
    assert len(calls) == 1 and calls[0] >= 10               # one batch for the whole list
