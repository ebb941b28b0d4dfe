"""Regenerates tests/golden/locate_bp_reference.json by running the REFERENCE'S OWN stage function
nanomonsv/locate_bp.py:88-109 (`locate_bp`, with get_refined_bp :13-84) from /root/reference, unmodified, TOGETHER WITH THE
REFERENCE'S OWN smith_waterman.sw_jump (pure numpy, imported as it is): nothing of this repository takes part in the numbers.
Only `pysam` is replaced (FastaFile over in-memory contigs, registered in sys.modules before the import). Inputs and produced
outputs only are stored.

    python tests/golden/make_golden_locate_bp.py          (build container, /root/reference present; about a minute:
                                                           the reference's sw_jump is a Python double loop)
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nanomonsv_b200.my_seq import reverse_complement    # noqa: E402  (only to build the synthetic contigs)

CONTIGS = {}


class _FastaFile:
    def __init__(self, path):
        pass

    def fetch(self, chrom, start, end):
        return CONTIGS[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(CONTIGS[chrom])

    def close(self):
        pass


def load():
    pysam = types.ModuleType("pysam")
    pysam.FastaFile = _FastaFile
    sys.modules["pysam"] = pysam
    pkg = types.ModuleType("nanomonsv")
    pkg.__path__ = ["/root/reference/nanomonsv"]
    sys.modules["nanomonsv"] = pkg
    return importlib.import_module("nanomonsv.locate_bp")


def rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))


def mutate(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append("ACGT"[rng.integers(0, 4)]); continue
        out.append(ch)
        if r > 1 - rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])
    return "".join(out)


def junction(rng, c1, bp1, d1, c2, bp2, d2, ins, L, rate=0.01):
    """Consensus contig of a junction: L bases ending at bp1 as read along dir1, inserted bases, L bases from bp2 on along dir2."""
    r1, r2 = CONTIGS[c1].upper(), CONTIGS[c2].upper()
    left = r1[max(bp1 - L, 0):bp1] if d1 == '+' else reverse_complement(r1[bp1 - 1:bp1 - 1 + L])
    right = r2[bp2 - 1:bp2 - 1 + L] if d2 == '-' else reverse_complement(r2[max(bp2 - L, 0):bp2])
    return mutate(rng, left, rate) + ins + mutate(rng, right, rate)


def main():
    rng = np.random.default_rng(20261021)
    CONTIGS["chrA"] = rand_seq(rng, 9000)
    b = rand_seq(rng, 9000)
    CONTIGS["chrB"] = b[:4000] + b[4000:4200].lower() + b[4200:]
    lines = []

    def add(c1, s1, e1, d1, c2, s2, e2, d2, mode, contig):
        lines.append(f"{c1},{s1},{e1},{d1},{c2},{s2},{e2},{d2},{mode},{len(lines)}\t{contig}")
    # deletions / duplications / translocations / inversions: breakpoint ranges of a few bases around the true positions
    add("chrA", 2995, 3004, '+', "chrA", 5197, 5203, '-', 'd', junction(rng, "chrA", 3000, '+', "chrA", 5200, '-', "", 220))
    add("chrA", 2990, 3010, '+', "chrA", 5190, 5210, '-', 'd', junction(rng, "chrA", 3000, '+', "chrA", 5200, '-', "GATTACA", 260))
    add("chrA", 6995, 7005, '-', "chrA", 6395, 6405, '+', 'r', junction(rng, "chrA", 7000, '-', "chrA", 6400, '+', "", 240))
    add("chrA", 1495, 1505, '+', "chrB", 4095, 4105, '-', 't', junction(rng, "chrA", 1500, '+', "chrB", 4100, '-', rand_seq(rng, 60), 300))
    add("chrA", 1495, 1505, '-', "chrB", 6095, 6105, '-', 't', junction(rng, "chrA", 1500, '-', "chrB", 6100, '-', "ACGGT", 250))
    add("chrA", 4095, 4105, '+', "chrB", 2495, 2505, '+', 't', junction(rng, "chrA", 4100, '+', "chrB", 2500, '+', "", 250))
    add("chrA", 3, 12, '-', "chrB", 8990, 9005, '+', 't', junction(rng, "chrA", 8, '-', "chrB", 8995, '+', "", 200))      # regions clipped at both contig ends
    add("chrA", 3, 12, '+', "chrB", 8990, 9005, '-', 't', junction(rng, "chrA", 8, '+', "chrB", 8995, '-', "", 200))      # 8 + 6 contig bases: below the minimum score
    add("chrA", 2995, 3004, '+', "chrA", 5197, 5203, '-', 'd', rand_seq(rng, 400))                                         # unrelated contig: None
    add("chrA", 0, 6, '-', "chrA", 5197, 5203, '-', 'd', junction(rng, "chrA", 5, '-', "chrA", 5200, '-', "TT", 150))     # start 0 -> clamped to 1
    # insertions: the reference window is the first sequence, the contig's ends are the regions
    ref = CONTIGS["chrA"].upper()
    add("chrA", 7490, 7500, '+', "chrA", 7501, 7512, '-', 'i', mutate(rng, ref[7300:7500], 0.01) + rand_seq(rng, 320) + mutate(rng, ref[7500:7700], 0.01))
    add("chrA", 800, 810, '+', "chrA", 811, 822, '-', 'i', mutate(rng, ref[400:805], 0.01) + rand_seq(rng, 900) + mutate(rng, ref[805:1250], 0.01))   # contig > 1000: 500-base ends
    add("chrA", 60, 70, '+', "chrA", 71, 80, '-', 'i', mutate(rng, ref[0:65], 0.0) + rand_seq(rng, 150) + mutate(rng, ref[65:200], 0.0))               # window clipped at 0
    mod = load()
    with tempfile.TemporaryDirectory() as tmp:
        cfile, out = os.path.join(tmp, "consensus.txt"), os.path.join(tmp, "out.txt")
        open(cfile, "w").write("\n".join(lines) + "\n")
        mod.locate_bp(cfile, out, "unused.fa", (1, 3, 3, 2), True)
        fix = {"generator": "tests/golden/make_golden_locate_bp.py",
               "reference": "nanomonsv v0.8.0 (/root/reference): locate_bp.py and smith_waterman.py imported unmodified; pysam.FastaFile stand-in",
               "sw_jump_params": [1, 3, 3, 2], "contigs": dict(CONTIGS), "consensus_lines": lines,
               "output": open(out).read().splitlines(), "log": open(out + ".locate_bp.log").read().splitlines()}
    json.dump(fix, open(os.path.join(HERE, "locate_bp_reference.json"), "w"), indent=0)
    print(len(fix["output"]), "of", len(lines), "contigs located")
    for l in fix["output"]:
        print(l[:100])


if __name__ == "__main__":
    main()
