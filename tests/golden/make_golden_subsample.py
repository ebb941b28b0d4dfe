"""Regenerates tests/golden/subsample_vectors.json with the REFERENCE'S OWN bam_subsample_fetch
(nanomonsv/count_sread_by_alignment.py:11-23): which reads of a +-100 bp window survive the (unseeded in production)
random subsampling to 500 reads. The function is taken out of the source with `ast` (the module imports pysam and
parasail) and run with a seeded `random` on stand-in alignment handles of 1, 499, 500, 501 and 1300 reads.
Nothing of the reference's source is written into this repository: inputs and produced outputs only.

    python tests/golden/make_golden_subsample.py        (build container, /root/reference present)
"""
import ast
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nanomonsv/count_sread_by_alignment.py"


def load_reference_function(name):
    tree = ast.parse(open(SRC).read(), SRC)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name]
    assert len(fn) == 1, name
    ns = {"random": random}
    exec(compile(ast.Module(body=fn, type_ignores=[]), SRC, "exec"), ns)
    return ns[name]


class Handle:
    def __init__(self, n):
        self.n = n

    def fetch(self, chrom, start, end):
        return iter(range(self.n))


def main():
    ref = load_reference_function("bam_subsample_fetch")
    cases = []
    for n, seed, cap in ((1, 1, 500), (499, 2, 500), (500, 3, 500), (501, 4, 500), (1300, 5, 500), (40, 6, 7)):
        random.seed(seed)
        kept = list(ref(Handle(n), "chr1", 0, 200, cap))
        cases.append({"n": n, "seed": seed, "max_read_num": cap, "kept": kept})
        print(n, cap, len(kept))
    json.dump({"source": "nanomonsv/count_sread_by_alignment.py:11-23 run from /root/reference (tests/golden/make_golden_subsample.py)",
               "cases": cases}, open(os.path.join(HERE, "subsample_vectors.json"), "w"))


if __name__ == "__main__":
    main()
