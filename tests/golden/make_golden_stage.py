"""Regenerates tests/golden/stage_reference.json by running the REFERENCE'S OWN stage function
nanomonsv/count_sread_by_alignment.py:401 (`count_sread_by_alignment`: read gathering, GNU sort, template
construction, the two ssw_check strands, decision rules, file formats) from /root/reference, unmodified, on a small
synthetic tumor sample with canonical and single-breakend candidates.

Two third-party modules that the reference imports are absent in this image and are replaced by stand-ins that are
registered in sys.modules before the import:
  * `pysam`    -> AlignmentFile / FastaFile objects over in-memory records and contigs (the records and contigs are
                  stored in the fixture, so the run can be repeated with a real pysam);
  * `parasail` -> matrix_create / ssw whose numbers come from THIS repository's CPU oracle (oracle/nmsv_oracle.c).
So the fixture pins every line of the reference's host logic on this path -- gather rule, mapq bookkeeping, template
quirks, strand rule and coordinate mirroring of ssw_check, thresholds, margins, mapq filter, output formats -- on the
reference itself; only the arithmetic inside parasail.ssw remains a restatement.
Nothing of the reference's source is written into this repository: inputs and produced outputs only.

    python tests/golden/make_golden_stage.py          (build container, /root/reference present)
"""
import dataclasses
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nanomonsv_b200 import synth                       # noqa: E402
from nanomonsv_b200.my_seq import reverse_complement    # noqa: E402
from oracle import oracle as O                          # noqa: E402

REGISTRY = {"records": [], "contigs": {}}


# ---- stand-in for pysam ---------------------------------------------------------------------------------------------
class _Read:
    def __init__(self, d):
        self.qname, self.flag, self.mapping_quality, self.query_sequence = d["qname"], d["flag"], d["mapq"], d["seq"]
        self._chrom, self._start, self._end = d["chrom"], d["start"], d["end"]


class _AlignmentFile:
    def __init__(self, path, mode="rb", **kw):
        self._reads = [_Read(d) for d in REGISTRY["records"]]

    def fetch(self, contig=None, start=None, stop=None):
        for r in self._reads:
            if contig is None or (r._chrom == contig and r._start < stop and r._end > start):
                yield r

    def close(self):
        pass


class _FastaFile:
    def __init__(self, path):
        pass

    def fetch(self, chrom, start, end):
        s = REGISTRY["contigs"][chrom]
        return s[max(start, 0):max(end, 0)]

    def close(self):
        pass


# ---- stand-in for parasail (numbers from the oracle) ---------------------------------------------------------------------
class _Matrix:
    def __init__(self, alphabet, match, mismatch):
        assert alphabet == "ACGT"
        self.match, self.mismatch = match, mismatch


def _ssw(s1, s2, gap_open, gap_extend, matrix):
    return O.ssw(s1, s2, gap_open, gap_extend, matrix.match, matrix.mismatch)      # fields score1, read_begin1, ... ; None on saturation


def load_reference(real_parasail=None):
    """real_parasail: an imported parasail module to use instead of the oracle-backed stand-in (tests/test_vs_parasail.py
    re-runs the reference with the real engine wherever it is installed)."""
    pysam = types.ModuleType("pysam")
    pysam.AlignmentFile, pysam.FastaFile = _AlignmentFile, _FastaFile
    parasail = real_parasail
    if parasail is None:
        parasail = types.ModuleType("parasail")
        parasail.matrix_create, parasail.ssw = (lambda a, m, x: _Matrix(a, m, x)), _ssw
    sys.modules["pysam"], sys.modules["parasail"] = pysam, parasail
    pkg = types.ModuleType("nanomonsv")                 # the package without its __init__ (which would pull in the CLI)
    pkg.__path__ = ["/root/reference/nanomonsv"]
    sys.modules["nanomonsv"] = pkg
    import importlib
    return importlib.import_module("nanomonsv.count_sread_by_alignment")


def records_of(wl, sample, rng, sbnd):
    """BAM-like records: primary record at one breakpoint, supplementary at the other, reverse-strand storage, noise."""
    recs = []
    for sv, reads in zip(wl.svs, wl.samples[sample]):
        f = sv.key.split(",")
        bp1, bp2 = (f[0], int(f[1])), (None if sbnd else (f[3], int(f[4])))
        for r in reads:
            seq = synth.codes_to_str(r.codes)
            rev = bool(rng.random() < 0.5)
            stored = reverse_complement(seq) if rev else seq
            at1, at2 = r.mapq1 is not None, (not sbnd) and r.mapq2 is not None
            (c, p), q = (bp1, r.mapq1) if at1 else (bp2, r.mapq2)
            recs.append(dict(qname=r.rid, chrom=c, start=p - 60, end=p + 60, mapq=int(q), flag=16 if rev else 0, seq=stored))
            if at1 and at2:
                recs.append(dict(qname=r.rid, chrom=bp2[0], start=bp2[1] - 60, end=bp2[1] + 60, mapq=int(r.mapq2),
                                 flag=2048 | (16 if rev else 0), seq=stored[:50]))
    f = wl.svs[0].key.split(",")
    recs.append(dict(qname="dupread", chrom=f[0], start=int(f[1]) - 10, end=int(f[1]) + 10, mapq=60, flag=1024, seq="ACGT" * 100))
    recs.append(dict(qname="secread", chrom=f[0], start=int(f[1]) - 10, end=int(f[1]) + 10, mapq=60, flag=256, seq="ACGT" * 100))
    recs.append(dict(qname="faraway", chrom=f[0], start=5, end=300, mapq=60, flag=0, seq="ACGT" * 100))
    return recs


def build_inputs(seed, rng):
    common = dict(coverage=6.0, read_len_mean=1300, read_len_min=500, read_len_max=2600, contig_len=120000, n_contigs=2, seed=seed)
    wl = synth.make_workload(dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=9, short_fraction=0.5, **common),
                             samples=("tumor",))
    swl = synth.make_workload(dataclasses.replace(synth.PRESETS["single_breakend"], n_sv=5, sbnd_contig_range=(100, 600),
                                                  sbnd_tail_range=(1200, 2500), **common), samples=("tumor",))
    assert all((swl.fasta.contigs[k] == wl.fasta.contigs[k]).all() for k in wl.fasta.contigs)
    contigs = {k: synth.codes_to_str(v) for k, v in wl.fasta.contigs.items()}
    # soft-masked stretch inside one template window and an N run: the .upper() and wildcard paths of the reference
    k0 = wl.svs[0].key.split(",")
    c = contigs[k0[0]]
    p = int(k0[1])
    contigs[k0[0]] = c[:p - 150] + c[p - 150:p - 120].lower() + c[p - 120:p + 40] + "N" * 6 + c[p + 46:]
    recs = records_of(wl, "tumor", rng, False) + records_of(swl, "tumor", rng, True)
    order = rng.permutation(len(recs))
    return wl, swl, contigs, [recs[i] for i in order]


def live_short(f):
    ins = 0 if f[6] == "---" else len(f[6])
    return f[0] == f[3] and (f[2], f[5]) in (("+", "-"), ("-", "+")) and int(f[4]) - int(f[1]) + len(str(ins)) < 300


def main():
    ref = load_reference()
    params = dict(var_read_min_mapq=40, score_ratio_thres=1.2, sort_option="-S 64M")
    for seed in range(77, 400):
        # a seed whose candidates hold an insertion, short del/dup candidates of both orientations and long ones ...
        rng = np.random.default_rng(2026 + seed)
        wl, swl, contigs, recs = build_inputs(seed, rng)
        ks = [s_.key.split(",") for s_ in wl.svs]
        if not (any(f[6] != "---" for f in ks) and {(f[2], f[5]) for f in ks if live_short(f)} == {("+", "-"), ("-", "+")}
                and sum(not live_short(f) for f in ks) >= 3 and {s_.key.split(",")[2] for s_ in swl.svs} == {"+", "-"}):
            continue
        REGISTRY["records"], REGISTRY["contigs"] = recs, contigs
        sv_lines = ["Chr_1\tPos_1\tDir_1\tChr_2\tPos_2\tDir_2\tInserted_Seq\tSV_ID"] + ["\t".join(s_.key.split(",")) for s_ in wl.svs]
        sbnd_lines = ["\t".join(s_.key.split(",")) for s_ in swl.svs]
        with tempfile.TemporaryDirectory() as tmp:
            fsv, fsb = os.path.join(tmp, "refined_bp.txt"), os.path.join(tmp, "refined_bp.sbnd.txt")
            open(fsv, "w").write("".join(l + "\n" for l in sv_lines))
            open(fsb, "w").write("".join(l + "\n" for l in sbnd_lines))
            out = {k: os.path.join(tmp, k) for k in ("count", "info", "count_sbnd", "info_sbnd")}
            ref.count_sread_by_alignment(fsv, os.path.join(tmp, "tumor.bam"), out["count"], out["info"], os.path.join(tmp, "ref.fa"),
                                         sbnd_file=fsb, output_count_file_sbnd=out["count_sbnd"],
                                         output_alignment_info_file_sbnd=out["info_sbnd"], debug=True, **params)
            import gc
            gc.collect()                                 # Alignment_counter.__del__ closes (flushes) the output files
            produced = {k: open(v).read().splitlines() for k, v in out.items()}
            produced["seam"] = open(out["count"] + ".tmp.local_read_for_realignment").read().splitlines()
            produced["seam_sbnd"] = open(out["count_sbnd"] + ".tmp.local_read_for_realignment").read().splitlines()
        # ... and whose run has supporting reads for insertion, short and long candidates and on the single-breakend pass
        info_keys = {tuple(l.split("\t")[:8]) for l in produced["info"]}
        if (len(produced["info_sbnd"]) >= 3 and any(k[6] != "---" for k in info_keys)
                and any(live_short(list(k)) for k in info_keys) and any(not live_short(list(k)) for k in info_keys)):
            break
    json.dump({"source": "nanomonsv/count_sread_by_alignment.py:401 run from /root/reference with stand-in pysam / parasail "
                         "(tests/golden/make_golden_stage.py)", "seed": seed,
               "contigs": contigs, "records": recs, "sv_file": sv_lines, "sbnd_file": sbnd_lines, "params": params,
               "expected": produced}, open(os.path.join(HERE, "stage_reference.json"), "w"), indent=0)
    print("seed", seed, {k: len(v) for k, v in produced.items()})


if __name__ == "__main__":
    main()
