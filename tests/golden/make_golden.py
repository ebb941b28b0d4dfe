"""Regenerates the machine-made golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

  ssw_fuzz_vectors.json   400 seeded random / tie-heavy / wildcard cases; expected values come from the
                          pure-Python full-matrix restatement oracle.ssw_pure_python,
                          i.e. from a formulation independent of both the C oracle and the CUDA kernel.
  validate_small.json     a small synthetic validation workload (seam TSV rows + reference contigs) with the
                          sread_count / sread_info lines produced by the scalar C oracle's restatement of
                          Alignment_counter (count_sread_by_alignment.py:207-397).

The reference itself cannot be imported here (pysam and parasail are absent, no network), so no fixture is
produced by the reference: PARITY UNPINNED at the parasail boundary.
"""
import dataclasses
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from nanomonsv_b200 import synth  # noqa: E402


def fuzz_vectors(n=400, seed=7):
    rnd = random.Random(seed)
    out = []
    for it in range(n):
        alpha = rnd.choice(["AC", "ACGT", "ACGTN", "AAAC", "ACGTacgtNX"])
        m = rnd.randint(1, 60)
        a = "".join(rnd.choice(alpha) for _ in range(m))
        if rnd.random() < 0.6:
            b = list(a if rnd.random() < 0.5 else O.reverse_complement(a))
            for _ in range(rnd.randint(0, 6)):
                if not b:
                    break
                p = rnd.randrange(len(b))
                op = rnd.random()
                if op < 0.4:
                    b[p] = rnd.choice(alpha)
                elif op < 0.7:
                    del b[p]
                else:
                    b.insert(p, rnd.choice(alpha))
            b = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 30))) + "".join(b) + \
                "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 30)))
            b = b or "A"
        else:
            b = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 100)))
        sc = rnd.choice([(2, -2, 3, 1), (2, -2, 3, 1), (1, -2, 3, 1), (2, -3, 5, 2), (3, -1, 2, 2)])
        out.append({"s1": a, "s2": b, "match": sc[0], "mismatch": sc[1], "open": sc[2], "extend": sc[3],
                    "expected": list(O.ssw_pure_python(a, b, sc[2], sc[3], sc[0], sc[1]))})
    return out


def validate_fixture():
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=8, coverage=6.0, read_len_mean=1500,
                              read_len_min=600, read_len_max=3000, contig_len=60000, n_contigs=2, seed=424242,
                              short_fraction=0.5)
    wl = synth.make_workload(cfg)
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    fix = {"contigs": {k: synth.codes_to_str(v) for k, v in wl.fasta.contigs.items()}, "samples": {}}
    for sample in ("tumor", "control"):
        seam = os.path.join(tmp, sample + ".seam.tsv")
        wl.write_seam(sample, seam)
        O.count_from_seam(seam, wl.fasta, seam + ".count", seam + ".info", var_read_min_mapq=40)
        fix["samples"][sample] = {"seam": open(seam).read().splitlines(),
                                  "count": open(seam + ".count").read().splitlines(),
                                  "info": sorted(open(seam + ".info").read().splitlines())}
    return fix


if __name__ == "__main__":
    with open(os.path.join(HERE, "ssw_fuzz_vectors.json"), "w") as h:
        json.dump({"vectors": fuzz_vectors()}, h, indent=0)
    with open(os.path.join(HERE, "validate_small.json"), "w") as h:
        json.dump(validate_fixture(), h, indent=0)
    print("golden fixtures written")
