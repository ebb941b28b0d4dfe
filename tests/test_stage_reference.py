"""The stage seam against the REFERENCE'S OWN run of count_sread_by_alignment (tests/golden/stage_reference.json, made by
tests/golden/make_golden_stage.py: the unmodified reference module executed from /root/reference with in-memory
stand-ins for pysam and with parasail's numbers supplied by the oracle).

  * not gpu: the host mirror's read gathering must reproduce the reference's two sorted seam files line for line, and the
    oracle's restatement of Alignment_counter must reproduce the four count / info files from those seam files;
  * gpu:     the drop-in stage function (gather, GNU sort, templates, CUDA validation, writers) must reproduce the four
    files from the same records.
Info lines are compared as sorted lists: within one SV the reference writes them in set-iteration order (:325).
"""
import dataclasses
import json
import os

import pytest

from nanomonsv_b200 import count_sread_by_alignment as C
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stage_reference.json")


@dataclasses.dataclass
class Rec:
    qname: str
    reference_name: str
    reference_start: int
    reference_end: int
    mapping_quality: int
    query_sequence: str
    flag: int

    is_reverse = property(lambda s: bool(s.flag & 16))
    is_secondary = property(lambda s: bool(s.flag & 256))
    is_duplicate = property(lambda s: bool(s.flag & 1024))
    is_supplementary = property(lambda s: bool(s.flag & 2048))


class Records:
    """What the host code needs from pysam.AlignmentFile: fetch(contig, start, stop), fetch(), close()."""

    def __init__(self, recs):
        self.recs = [Rec(d["qname"], d["chrom"], d["start"], d["end"], d["mapq"], d["seq"], d["flag"]) for d in recs]

    def fetch(self, contig=None, start=None, stop=None):
        for r in self.recs:
            if contig is None or (r.reference_name == contig and r.reference_start < stop and r.reference_end > start):
                yield r

    def close(self):
        pass


@pytest.fixture(scope="module")
def fix():
    return json.load(open(GOLDEN))


def _write_inputs(fix, tmp_path):
    sv_file, sbnd_file = tmp_path / "refined_bp.txt", tmp_path / "refined_bp.sbnd.txt"
    sv_file.write_text("".join(l + "\n" for l in fix["sv_file"]))
    sbnd_file.write_text("".join(l + "\n" for l in fix["sbnd_file"]))
    return str(sv_file), str(sbnd_file)


def test_gather_and_oracle_reproduce_the_reference_run(fix, tmp_path, monkeypatch):
    exp = fix["expected"]
    sv_file, sbnd_file = _write_inputs(fix, tmp_path)
    monkeypatch.setattr(C, "_get_alignment_object", lambda a, r: Records(fix["records"]))
    seam, seam_sbnd = str(tmp_path / "seam"), str(tmp_path / "seam_sbnd")
    C.gather_local_read_for_realignment(sv_file, "tumor.bam", seam, "ref.fa", sbnd_file=sbnd_file, output_file_sbnd=seam_sbnd,
                                        sort_option=fix["params"]["sort_option"])
    assert open(seam).read().splitlines() == exp["seam"]
    assert open(seam_sbnd).read().splitlines() == exp["seam_sbnd"]
    assert len(exp["seam"]) > 100 and len(exp["seam_sbnd"]) > 40

    fasta = O.DictFasta(fix["contigs"])
    q, ratio = fix["params"]["var_read_min_mapq"], fix["params"]["score_ratio_thres"]
    O.count_from_seam(seam, fasta, seam + ".count", seam + ".info", var_read_min_mapq=q, score_ratio_thres=ratio)
    O.count_from_seam(seam_sbnd, fasta, seam_sbnd + ".count", seam_sbnd + ".info", is_sbnd=True, var_read_min_mapq=q,
                      score_ratio_thres=ratio)
    assert open(seam + ".count").read().splitlines() == exp["count"]
    assert sorted(open(seam + ".info").read().splitlines()) == sorted(exp["info"])
    assert open(seam_sbnd + ".count").read().splitlines() == exp["count_sbnd"]
    assert sorted(open(seam_sbnd + ".info").read().splitlines()) == sorted(exp["info_sbnd"])
    # the run exercises what it is meant to pin: supporting reads on both passes, a short del/dup (33-column info
    # lines with reference-segment tuples) and a long one (12 x "None"), both strands
    assert len(exp["info"]) >= 5 and len(exp["info_sbnd"]) >= 3
    cols = [l.split("\t") for l in exp["info"]]
    assert any(c[-1] == "None" for c in cols) and any(c[-1] in "+-" for c in cols)
    assert {c[14] for c in cols} == {"+", "-"}


@pytest.mark.gpu
def test_stage_function_reproduces_the_reference_run(fix, tmp_path, monkeypatch, engine):
    exp = fix["expected"]
    sv_file, sbnd_file = _write_inputs(fix, tmp_path)
    monkeypatch.setattr(C, "_get_alignment_object", lambda a, r: Records(fix["records"]))
    out = {k: str(tmp_path / k) for k in ("count", "info", "count_sbnd", "info_sbnd")}
    C.count_sread_by_alignment(sv_file, "tumor.bam", out["count"], out["info"], O.DictFasta(fix["contigs"]),
                               sbnd_file=sbnd_file, output_count_file_sbnd=out["count_sbnd"],
                               output_alignment_info_file_sbnd=out["info_sbnd"], debug=True, engine=engine, **fix["params"])
    assert open(out["count"]).read().splitlines() == exp["count"]
    assert sorted(open(out["info"]).read().splitlines()) == sorted(exp["info"])
    assert open(out["count_sbnd"]).read().splitlines() == exp["count_sbnd"]
    assert sorted(open(out["info_sbnd"]).read().splitlines()) == sorted(exp["info_sbnd"])
    assert open(out["count"] + ".tmp.local_read_for_realignment").read().splitlines() == exp["seam"]


class _RecordingEngine:
    def __init__(self):
        self.batches = []

    def validate(self, pack, svs, reads, scoring=None):
        import numpy as np
        from nanomonsv_b200 import _lib
        codes, offsets, lens = pack.finalize()
        self.batches.append((codes.copy(), offsets.copy(), lens.copy(), svs.copy(), reads.copy()))
        checked = np.array([int(e) - int(b) for b, e in zip(svs["read_begin"], svs["read_end"])], dtype=np.int32)
        return checked, np.zeros(len(svs), dtype=np.int32), np.zeros(0, dtype=_lib.RECORD_DTYPE)


def _per_sv_reads(batches):
    """[(templates as code bytes, {(mapq1, mapq2, read as code bytes)})] over all batches, in SV order."""
    out = []
    for codes, offsets, lens, svs, reads in batches:
        def seq(i):      # a read that is only counted (mapq test fails) travels as a length without a sequence
            if int(offsets[i]) == 0xFFFFFFFFFFFFFFFF:
                return ("absent", int(lens[i]))
            return bytes(codes[int(offsets[i]):int(offsets[i]) + int(lens[i])])
        for sv in svs:
            tm = tuple(seq(int(t)) if int(t) != 0xffffffff else None for t in sv["tmpl"])
            rs = {(int(r["mapq1"]), int(r["mapq2"]), seq(int(r["seq"]))) for r in reads[int(sv["read_begin"]):int(sv["read_end"])]}
            out.append((tm, rs))
    return out


def test_direct_gather_builds_the_same_work(fix, tmp_path, monkeypatch):
    """NMSV_DIRECT_GATHER=1 (reads kept in memory, no seam file, no sort of whole reads) hands the GPU the same SV
    candidates in the same order with the same reads, mapqs and templates as the seam-file path -- and packs a read that
    serves several candidates once."""
    sv_file, sbnd_file = _write_inputs(fix, tmp_path)
    monkeypatch.setattr(C, "_get_alignment_object", lambda a, r: Records(fix["records"]))
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NMSV_DIRECT_GATHER", mode)
        eng = _RecordingEngine()
        out = {k: str(tmp_path / (k + mode)) for k in ("count", "info", "count_sbnd", "info_sbnd")}
        C.count_sread_by_alignment(sv_file, "tumor.bam", out["count"], out["info"], O.DictFasta(fix["contigs"]),
                                   sbnd_file=sbnd_file, output_count_file_sbnd=out["count_sbnd"],
                                   output_alignment_info_file_sbnd=out["info_sbnd"], engine=eng, **fix["params"])
        res[mode] = (eng.batches, open(out["count"]).read(), open(out["count_sbnd"]).read())
        assert not os.path.exists(out["count"] + ".tmp.local_read_for_realignment")
    assert res["0"][1] == res["1"][1] and res["0"][2] == res["1"][2]              # same candidates, same order, same read counts
    assert _per_sv_reads(res["0"][0]) == _per_sv_reads(res["1"][0])
    n_seq = {m: sum(len(b[2]) for b in res[m][0]) for m in res}
    assert n_seq["1"] <= n_seq["0"]


@pytest.mark.gpu
def test_direct_gather_reproduces_the_reference_run(fix, tmp_path, monkeypatch, engine):
    exp = fix["expected"]
    sv_file, sbnd_file = _write_inputs(fix, tmp_path)
    monkeypatch.setattr(C, "_get_alignment_object", lambda a, r: Records(fix["records"]))
    monkeypatch.setenv("NMSV_DIRECT_GATHER", "1")
    out = {k: str(tmp_path / k) for k in ("count", "info", "count_sbnd", "info_sbnd")}
    C.count_sread_by_alignment(sv_file, "tumor.bam", out["count"], out["info"], O.DictFasta(fix["contigs"]),
                               sbnd_file=sbnd_file, output_count_file_sbnd=out["count_sbnd"],
                               output_alignment_info_file_sbnd=out["info_sbnd"], engine=engine, **fix["params"])
    assert open(out["count"]).read().splitlines() == exp["count"]
    assert sorted(open(out["info"]).read().splitlines()) == sorted(exp["info"])
    assert open(out["count_sbnd"]).read().splitlines() == exp["count_sbnd"]
    assert sorted(open(out["info_sbnd"]).read().splitlines()) == sorted(exp["info_sbnd"])
