"""Regenerates tests/golden/legacy_reference.json by running the REFERENCE'S OWN legacy entry point
nanomonsv/long_read_validate.py:421 (`long_read_validate_main`) from /root/reference, unmodified, on a small synthetic
tumor / control pair -- same stand-ins for the absent pysam and parasail modules as make_golden_stage.py (records and
contigs stored in the fixture; parasail's numbers from this repository's oracle). The contigs are partly soft-masked:
the legacy module does not upper-case its templates, so lower-case bases reach the aligner (wildcards) and
my_seq.reverse_complement leaves them uncomplemented.

    python tests/golden/make_golden_legacy.py          (build container, /root/reference present)
"""
import dataclasses
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_stage as S      # noqa: E402  (stand-in modules, record builder)

from nanomonsv_b200 import synth   # noqa: E402


def main():
    S.load_reference()
    import importlib
    ref = importlib.import_module("nanomonsv.long_read_validate")
    rng = np.random.default_rng(2027)
    def legacy_short(key):
        f = key.split(",")
        ins = 0 if f[6] == "---" else len(f[6])
        return f[0] == f[3] and (f[2], f[5]) in (("+", "-"), ("-", "+")) and int(f[4]) - int(f[1]) + len(str(ins)) < 100

    for seed in range(91, 400):       # first seed whose candidates reach the legacy short-del/dup branch in both orientations
        cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=9, short_fraction=0.7, coverage=7.0,
                                  read_len_mean=1300, read_len_min=500, read_len_max=2600, contig_len=120000, n_contigs=2,
                                  seed=seed)
        wl = synth.make_workload(cfg)
        shorts = [s.key.split(",") for s in wl.svs if legacy_short(s.key)]
        if len(shorts) >= 3 and {(f[2], f[5]) for f in shorts} == {("+", "-"), ("-", "+")}:
            break
    print("seed", seed, "legacy-short candidates", len(shorts))
    contigs = {}
    for k, v in wl.fasta.contigs.items():
        s = synth.codes_to_str(v)
        # soft-mask 23 of every 211 bases (like a repeat-masked genome)
        contigs[k] = "".join(c.lower() if (i % 211) < 23 else c for i, c in enumerate(s))
    recs = {smp: S.records_of(wl, smp, rng, False) for smp in ("tumor", "control")}
    for smp in recs:
        order = rng.permutation(len(recs[smp]))
        recs[smp] = [recs[smp][i] for i in order]
    sv_lines = ["Chr_1\tPos_1\tDir_1\tChr_2\tPos_2\tDir_2\tInserted_Seq\tSV_ID"] + ["\t".join(s.key.split(",")) for s in wl.svs]
    S.REGISTRY["contigs"] = contigs

    class _ByPath(S._AlignmentFile):          # tumor.bam / control.bam -> their own record lists
        def __init__(self, path, mode="rb", **kw):
            self._reads = [S._Read(d) for d in recs["control" if "control" in os.path.basename(path) else "tumor"]]
    sys.modules["pysam"].AlignmentFile = _ByPath

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fsv = os.path.join(tmp, "result.txt")
        open(fsv, "w").write("".join(l + "\n" for l in sv_lines))
        for name, control in (("tumor_only", None), ("tumor_control", os.path.join(tmp, "control.bam"))):
            fo, fs = os.path.join(tmp, name + ".out"), os.path.join(tmp, name + ".sread")
            ref.long_read_validate_main(fsv, os.path.join(tmp, "tumor.bam"), fo, fs, os.path.join(tmp, "ref.fa"), control, 40, False, False)
            out[name] = {"output": open(fo).read().splitlines(), "sread": open(fs).read().splitlines()}
    json.dump({"source": "nanomonsv/long_read_validate.py:421 run from /root/reference with stand-in pysam / parasail "
                         "(tests/golden/make_golden_legacy.py)",
               "contigs": contigs, "records": recs, "sv_file": sv_lines, "var_read_min_mapq": 40, "expected": out},
              open(os.path.join(HERE, "legacy_reference.json"), "w"), indent=0)
    print({k: (len(v["output"]), len(v["sread"])) for k, v in out.items()})
    print(out["tumor_control"]["output"][:4])


if __name__ == "__main__":
    main()
