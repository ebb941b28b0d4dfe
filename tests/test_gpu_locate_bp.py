"""SURVEY.md section 8f row 2, the stage function: locate_bp (reference locate_bp.py:88-109) with all consensus lines as one
sw_jump batch on the GPU, held to the output and log files that the reference's own locate_bp wrote with the reference's own
sw_jump (tests/golden/locate_bp_reference.json: nothing of this repository took part in those numbers)."""
import json
import os
import re

import pytest

from nanomonsv_b200 import locate_bp as LB

pytestmark = pytest.mark.gpu
FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "locate_bp_reference.json")))
# The fixture was written under numpy 2, whose scalars print as np.int32(195) inside a tuple; the reference's sw_jump returns numpy
# scalars, so its debug log carries that spelling (under numpy 1 it prints plain 195, which is what this repository writes).
LOG = [re.sub(r"np\.int(?:32|64)\((-?\d+)\)", r"\1", l) for l in FIX["log"]]


class _Fasta:
    def __init__(self, contigs):
        self.contigs = contigs

    def fetch(self, chrom, start, end):
        return self.contigs[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(self.contigs[chrom])


def test_locate_bp_equals_the_reference_run(engine, tmp_path):
    cfile, out = tmp_path / "consensus.txt", tmp_path / "out.txt"
    cfile.write_text("\n".join(FIX["consensus_lines"]) + "\n")
    fa = _Fasta(FIX["contigs"])
    LB.locate_bp(str(cfile), str(out), fa, tuple(FIX["sw_jump_params"]), True, engine=engine)
    assert out.read_text().splitlines() == FIX["output"]
    assert open(str(out) + ".locate_bp.log").read().splitlines() == LOG
    key, contig = FIX["consensus_lines"][3].split("\t")
    c1, s1, e1, d1, c2, s2, e2, d2, mode, cid = key.split(",")
    got = LB.get_refined_bp(contig, fa, c1, int(s1), int(e1), d1, c2, int(s2), int(e2), d2, mode, None, tuple(FIX["sw_jump_params"]), engine=engine)
    assert f"{c1}\t{got[0]}\t{d1}\t{c2}\t{got[1]}\t{d2}\t{got[2]}\t{mode}_{cid}" == FIX["output"][3]
