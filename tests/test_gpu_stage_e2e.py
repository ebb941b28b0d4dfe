"""End-to-end test of the stage seam: count_sread_by_alignment(sv_file, alignment_file, ...) with read gathering
(window fetches, mapq bookkeeping, primary-record pass, GNU sort), template construction, GPU validation and the four
output files -- pysam replaced by duck-typed objects (pysam is not installable in this image)."""
import dataclasses
import os

import numpy as np
import pytest

from nanomonsv_b200 import count_sread_by_alignment as C
from nanomonsv_b200 import synth
from nanomonsv_b200.my_seq import reverse_complement
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@dataclasses.dataclass
class Rec:
    qname: str
    reference_name: str
    reference_start: int
    reference_end: int
    mapping_quality: int
    query_sequence: str
    is_reverse: bool = False
    is_supplementary: bool = False
    is_secondary: bool = False
    is_duplicate: bool = False


class FakeAlignmentFile:
    def __init__(self, recs):
        self.recs = recs

    def fetch(self, chrom=None, start=None, end=None):
        for r in self.recs:
            if chrom is None or (r.reference_name == chrom and r.reference_start < end and r.reference_end > start):
                yield r

    def close(self):
        pass


def _records(wl, sample, rng, sbnd):
    recs = []
    for sv, reads in zip(wl.svs, wl.samples[sample]):
        f = sv.key.split(",")
        bp1 = (f[0], int(f[1]))
        bp2 = None if sbnd else (f[3], int(f[4]))
        for r in reads:
            seq = synth.codes_to_str(r.codes)
            rev = bool(rng.random() < 0.5)
            stored = reverse_complement(seq) if rev else seq          # a BAM stores the reference-strand sequence
            at1 = r.mapq1 is not None
            at2 = (not sbnd) and r.mapq2 is not None
            prim_bp, prim_q = (bp1, r.mapq1) if at1 else (bp2, r.mapq2)
            recs.append(Rec(r.rid, prim_bp[0], prim_bp[1] - 60, prim_bp[1] + 60, prim_q, stored, rev))
            if at1 and at2:                                            # split read: supplementary record at the other end
                recs.append(Rec(r.rid, bp2[0], bp2[1] - 60, bp2[1] + 60, r.mapq2, stored[:50], rev, is_supplementary=True))
    # noise the gather must ignore: a duplicate-flagged and a secondary record inside a window, a far-away read
    f = wl.svs[0].key.split(",")
    recs.append(Rec("dupread", f[0], int(f[1]) - 10, int(f[1]) + 10, 60, "ACGT" * 100, is_duplicate=True))
    recs.append(Rec("secread", f[0], int(f[1]) - 10, int(f[1]) + 10, 60, "ACGT" * 100, is_secondary=True))
    recs.append(Rec("faraway", f[0], 5, 300, 60, "ACGT" * 100))
    rng.shuffle(recs)
    return recs


def _expected_seam(keys, recs, sbnd):
    """Independent restatement of the reference's gathering rule (count_sread_by_alignment.py:35-120): per key the mapq
    of a read at a breakpoint is that of the LAST record of that qname overlapping the +-100 bp window (any flag); a line
    is written for every primary, non-duplicate record of a read seen in at least one window, with its sequence
    restored to the sequenced strand."""
    lines = set()
    for key in keys:
        f = key.split(",")
        wins = [(f[0], int(f[1]))] if sbnd else [(f[0], int(f[1])), (f[3], int(f[4]))]
        mapq = {}
        for w, (c, pos) in enumerate(wins):
            for r in recs:
                if r.reference_name == c and r.reference_start < pos + 100 and r.reference_end > max(pos - 100, 0):
                    mapq.setdefault(r.qname, [None, None])[w] = r.mapping_quality
        for r in recs:
            if r.qname in mapq and not (r.is_supplementary or r.is_secondary or r.is_duplicate):
                seq = reverse_complement(r.query_sequence) if r.is_reverse else r.query_sequence
                m1, m2 = mapq[r.qname]
                lines.add(f"{key}\t{r.qname}\t{m1}\t{'None' if sbnd else m2}\t{seq}")
    return lines


def test_count_sread_by_alignment_end_to_end(engine, tmp_path, monkeypatch):
    rng = np.random.default_rng(12)
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=7, coverage=6.0, read_len_mean=1500,
                              read_len_min=500, read_len_max=3000, contig_len=400000, n_contigs=2, seed=31,
                              short_fraction=0.5)
    wl = synth.make_workload(cfg, samples=("tumor",))
    scfg = dataclasses.replace(synth.PRESETS["single_breakend"], n_sv=5, coverage=6.0, read_len_mean=1500,
                               read_len_min=500, read_len_max=3000, contig_len=400000, n_contigs=2, seed=31,
                               sbnd_contig_range=(100, 600), sbnd_tail_range=(1500, 3000))
    swl = synth.make_workload(scfg, samples=("tumor",))
    swl.fasta = wl.fasta                       # one reference for both passes (same generator seed -> same contigs)
    assert all((swl.fasta.contigs[k] == wl.fasta.contigs[k]).all() for k in wl.fasta.contigs)
    # drop reads without any window (mapq1 and mapq2 both None cannot happen in the generator) and read-less SVs
    recs = _records(wl, "tumor", rng, False) + _records(swl, "tumor", rng, True)
    keys = list(dict.fromkeys(sv.key for sv in wl.svs))
    skeys = list(dict.fromkeys(sv.key for sv in swl.svs))
    monkeypatch.setattr(C, "_get_alignment_object", lambda a, r: FakeAlignmentFile(recs))

    sv_file, sbnd_file = tmp_path / "refined_bp.txt", tmp_path / "refined_bp.sbnd.txt"
    sv_file.write_text("Chr_1\tPos_1\tDir_1\tChr_2\tPos_2\tDir_2\tInserted_Seq\tSV_ID\n" +
                       "".join("\t".join(s.key.split(",")) + "\n" for s in wl.svs))
    sbnd_file.write_text("".join("\t".join(s.key.split(",")) + "\n" for s in swl.svs))
    out = {k: str(tmp_path / k) for k in ("count", "info", "count_sbnd", "info_sbnd")}
    C.count_sread_by_alignment(str(sv_file), "fake.bam", out["count"], out["info"], wl.fasta,
                               sbnd_file=str(sbnd_file), output_count_file_sbnd=out["count_sbnd"],
                               output_alignment_info_file_sbnd=out["info_sbnd"], var_read_min_mapq=40,
                               score_ratio_thres=1.2, sort_option="-S 64M", debug=True, engine=engine)
    # ---- the gathered, sorted seam files
    seam = out["count"] + ".tmp.local_read_for_realignment"
    seam_sbnd = out["count_sbnd"] + ".tmp.local_read_for_realignment"
    got = open(seam).read().splitlines()
    assert set(got) == _expected_seam(keys, recs, False) and len(got) == len(set(got))
    gk = [ln.split("\t")[0] for ln in got]
    assert all(gk[i] == gk[i - 1] or gk[i] not in gk[:i] for i in range(1, len(gk)))   # grouped by key
    assert set(open(seam_sbnd).read().splitlines()) == _expected_seam(skeys, recs, True)
    assert not any("faraway" in ln for ln in got)
    # duplicate / secondary records give a mapq but never a sequence line (:96-102)
    assert not any(ln.split("\t")[1] in ("dupread", "secread") for ln in got)
    # ---- counts and supporting-read lines against the oracle run on the same seam files
    O.count_from_seam(seam, wl.fasta, seam + ".o_count", seam + ".o_info", var_read_min_mapq=40, engine="striped")
    O.count_from_seam(seam_sbnd, wl.fasta, seam_sbnd + ".o_count", seam_sbnd + ".o_info", is_sbnd=True,
                      var_read_min_mapq=40, engine="striped")
    assert open(out["count"]).read() == open(seam + ".o_count").read()
    assert sorted(open(out["info"]).read().splitlines()) == sorted(open(seam + ".o_info").read().splitlines())
    assert open(out["count_sbnd"]).read() == open(seam_sbnd + ".o_count").read()
    assert sorted(open(out["info_sbnd"]).read().splitlines()) == sorted(open(seam_sbnd + ".o_info").read().splitlines())
    assert len(open(out["info"]).read().splitlines()) >= 3 and len(open(out["info_sbnd"]).read().splitlines()) >= 2
    # ---- without debug the intermediate files are removed, like the reference (:428-429, :450-451)
    C.count_sread_by_alignment(str(sv_file), "fake.bam", out["count"], out["info"], wl.fasta, var_read_min_mapq=40,
                               sort_option="-S 64M", debug=False, engine=engine)
    assert not os.path.exists(seam)
    # the count file feeds post_proc.integrate_realignment_result (post_proc.py:264-303): 10 tab-separated columns
    assert all(len(ln.split("\t")) == 10 for ln in open(out["count"]).read().splitlines())


def test_other_parasail_call_sites_through_the_facade(engine):
    """locate_bp_sbnd.py:28-29 uses matrix (1,-2); post_proc.py:118-120 compares inserted sequences with (2,-2)."""
    from nanomonsv_b200 import ssw as P
    import random
    rnd = random.Random(8)
    m12 = P.matrix_create("ACGT", 1, -2)
    m22 = P.matrix_create("ACGT", 2, -2)
    pairs = []
    for _ in range(30):
        a = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(50, 600)))
        b = a[rnd.randint(0, 20):] + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 200)))
        pairs.append((a, b))
    for m, sc in ((m12, (1, -2)), (m22, (2, -2))):
        got = P.ssw_batch(pairs, 3, 1, m, engine=engine)
        for (a, b), g in zip(pairs, got):
            e = O.ssw(a, b, 3, 1, sc[0], sc[1])
            assert (g.score1, g.read_begin1, g.read_end1, g.ref_begin1, g.ref_end1) == e.astuple()
