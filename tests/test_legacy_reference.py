"""The legacy module (nanomonsv/long_read_validate.py) against the REFERENCE'S OWN run of long_read_validate_main
(tests/golden/legacy_reference.json, made by tests/golden/make_golden_legacy.py with stand-in pysam / parasail modules):
tumor-only and tumor + control output files and the supporting-read file, on partly soft-masked contigs."""
import json
import os
import sys
import types

import pytest

from nanomonsv_b200 import long_read_validate as LRV

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "legacy_reference.json")


class _Read:
    def __init__(self, d):
        self.qname, self.flag, self.mapping_quality, self.query_sequence = d["qname"], d["flag"], d["mapq"], d["seq"]
        self.reference_name, self.reference_start, self.reference_end = d["chrom"], d["start"], d["end"]
        self.is_reverse, self.is_secondary = bool(d["flag"] & 16), bool(d["flag"] & 256)
        self.is_duplicate, self.is_supplementary = bool(d["flag"] & 1024), bool(d["flag"] & 2048)


def _fake_pysam(fix):
    class AlignmentFile:
        def __init__(self, path, mode="rb", **kw):
            self.reads = [_Read(d) for d in fix["records"]["control" if "control" in os.path.basename(path) else "tumor"]]

        def fetch(self, contig=None, start=None, stop=None):
            for r in self.reads:
                if contig is None or (r.reference_name == contig and r.reference_start < stop and r.reference_end > start):
                    yield r

        def close(self):
            pass

    class FastaFile:
        def __init__(self, path):
            pass

        def fetch(self, chrom, start, end):
            return fix["contigs"][chrom][max(start, 0):max(end, 0)]

        def close(self):
            pass

    mod = types.ModuleType("pysam")
    mod.AlignmentFile, mod.FastaFile = AlignmentFile, FastaFile
    return mod


def test_fixture_reaches_the_legacy_branches():
    fix = json.load(open(GOLDEN))
    keys = [LRV.legacy_key(l.split("\t")) for l in fix["sv_file"][1:]]
    assert sum(LRV.is_short_del_dup(k) for k in keys) >= 3           # the 4-template branch (ref1 tested twice)
    assert any(c.islower() for c in fix["contigs"]["chr1"])          # no upper-casing in the legacy module
    exp = fix["expected"]["tumor_control"]
    assert len(exp["output"]) == len(keys) and len(exp["sread"]) >= 10
    assert all(len(l.split("\t")) == 11 for l in exp["output"]) and all(len(l.split("\t")) == 9 for l in fix["expected"]["tumor_only"]["output"])


@pytest.mark.gpu
def test_legacy_main_reproduces_the_reference_run(tmp_path, monkeypatch, engine):
    fix = json.load(open(GOLDEN))
    monkeypatch.setitem(sys.modules, "pysam", _fake_pysam(fix))
    fsv = tmp_path / "result.txt"
    fsv.write_text("".join(l + "\n" for l in fix["sv_file"]))
    for name, control in (("tumor_only", None), ("tumor_control", str(tmp_path / "control.bam"))):
        fo, fs = str(tmp_path / (name + ".out")), str(tmp_path / (name + ".sread"))
        LRV.long_read_validate_main(str(fsv), str(tmp_path / "tumor.bam"), fo, fs, str(tmp_path / "ref.fa"), control,
                                    fix["var_read_min_mapq"], False, False, engine=engine)
        exp = fix["expected"][name]
        assert open(fo).read().splitlines() == exp["output"], name
        assert sorted(open(fs).read().splitlines()) == sorted(exp["sread"]), name
