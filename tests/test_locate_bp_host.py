"""Host logic of the locate_bp drop-in without a GPU: the sw_jump batch is answered by the C restatement of the reference's
function (tests only; itself pinned on the reference's own vectors), everything around it -- regions, clipping, strand
handling, positions, inserted sequence, output and log formats -- must reproduce the files that the reference's own locate_bp
wrote WITH THE REFERENCE'S OWN sw_jump (tests/golden/locate_bp_reference.json, generator make_golden_locate_bp.py).
The GPU twin is tests/test_gpu_locate_bp.py."""
import json
import os
import re

from nanomonsv_b200 import locate_bp as LB, smith_waterman as SW
from oracle import oracle as O

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


def test_locate_bp_host_logic(monkeypatch, tmp_path):
    def batch(problems, match_score=1, mismatch_penalty=2, gap_cost=2, jump_cost=2, minimum_score=10, engine=None):
        return [O.sw_jump(c, a, b, match_score, mismatch_penalty, gap_cost, jump_cost, minimum_score) for c, a, b in problems]
    monkeypatch.setattr(SW, "sw_jump_batch", batch)
    monkeypatch.setattr(SW, "sw_jump", lambda c, a, b, match_score=1, mismatch_penalty=2, gap_cost=2, jump_cost=2, minimum_score=10,
                        engine=None: batch([(c, a, b)], match_score, mismatch_penalty, gap_cost, jump_cost, minimum_score)[0])
    cfile, out = tmp_path / "consensus.txt", tmp_path / "out.txt"
    cfile.write_text("\n".join(FIX["consensus_lines"]) + "\n")
    fa = _Fasta(FIX["contigs"])
    LB.locate_bp(str(cfile), str(out), fa, tuple(FIX["sw_jump_params"]), True)
    assert out.read_text().splitlines() == FIX["output"]
    assert open(str(out) + ".locate_bp.log").read().splitlines() == LOG
    assert 0 < len(FIX["output"]) < len(FIX["consensus_lines"])            # located contigs and None cases are both present
    LB.locate_bp(str(cfile), str(out), fa, tuple(FIX["sw_jump_params"]), False)
    assert not os.path.exists(str(out) + ".locate_bp.log")
    # the per-contig function (:13-84)
    modes = set()
    for line in FIX["consensus_lines"]:
        key, contig = line.split("\t")
        c1, s1, e1, d1, c2, s2, e2, d2, mode, cid = key.split(",")
        got = LB.get_refined_bp(contig, fa, c1, int(s1), int(e1), d1, c2, int(s2), int(e2), d2, mode, None, tuple(FIX["sw_jump_params"]))
        exp = [l for l in FIX["output"] if l.endswith(f"\t{mode}_{cid}")]
        if got is None:
            assert exp == []
        else:
            assert exp == [f"{c1}\t{got[0]}\t{d1}\t{c2}\t{got[1]}\t{d2}\t{got[2]}\t{mode}_{cid}"]
            modes.add(mode)
    assert modes == {"d", "r", "t", "i"}
