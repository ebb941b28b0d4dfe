"""Regenerates tests/golden/sw_jump_vectors.json with the REFERENCE'S OWN implementation:
nanomonsv/smith_waterman.py (pure numpy) is loaded straight from /root/reference by file path (the package
__init__ would pull in pysam, which is absent). Run from the repo root in the build container:

    python tests/golden/make_golden_jump.py

The vectors pin the C oracle (oracle/nmsv_jump_oracle.c) and, on the GPU box, the CUDA path (nmsv_sw_jump_batch).
"""
import importlib.util
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_smith_waterman", "/root/reference/nanomonsv/smith_waterman.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def norm(r):
    if r is None:
        return None
    return [int(r[0]), [int(x) for x in r[1]], [int(x) for x in r[2]], [int(x) for x in r[3]], r[4], r[5]]


def cases(seed=20261019, n=120):
    rnd = random.Random(seed)

    def rs(k, alpha="ACGT"):
        return "".join(rnd.choice(alpha) for _ in range(k))

    out = []
    demo_a = ("AGGAGGGGGTTTCAGATTCCCCCATCCAAAGGAGGTTTTGGCCCCATGGGGAATGAAACAAAGACACATGAACACAGCTGCAAATATGCCCCACCTCTCCCTGCCCCCCCAAC"
              "TTGCCGGGCCAATGCACCTGCAGTCCCTCACCCTTCTCAGTTCATACCTACTCCCCAGCTCTCTCACTCTTGTATTCTAGATCCCTCAGTACACCAGCGCCCGGGAAGTCAGG"
              "GCCGGGCCAGGACTTTGACGTGCTGGTGGCAGACTGAAGGCCAACTCCCATGGGGAGTCCCCAAGGAGAAGGTGGAATCAGGGAAAAAAGGTGAATGGATGACTCAGTCACAG")
    out.append((demo_a, "CAAAGGGGAGTCTCCCAAGGAGAAGAGCCCAGGGCGCAAGGAGCAA", "TTACAATATTATCTGATGCTGAAAGGTGGGGATCAGGGAAAAG", (1, 2, 2, 2, 10)))
    for it in range(n):
        alpha = rnd.choice(["ACGT", "ACGT", "ACGT", "AC", "ACGTN"])
        r1, r2 = rs(rnd.randint(12, 90), alpha), rs(rnd.randint(12, 90), alpha)
        kind = it % 6
        if kind == 5:            # unrelated contig -> usually None
            c = rs(rnd.randint(20, 120), alpha)
        else:
            ins = rs(rnd.choice([0, 0, 0, 2, 9, 25]), alpha)
            c = rs(rnd.randint(0, 12), alpha) + r1[rnd.randint(0, 4):len(r1) - rnd.randint(0, 4)] + ins + \
                r2[rnd.randint(0, 4):len(r2) - rnd.randint(0, 4)] + rs(rnd.randint(0, 12), alpha)
            c = list(c)
            for _ in range(rnd.randint(0, 7)):
                p = rnd.randrange(len(c))
                op = rnd.random()
                if op < 0.4:
                    c[p] = rnd.choice(alpha)
                elif op < 0.7:
                    del c[p]
                else:
                    c.insert(p, rnd.choice(alpha))
            c = "".join(c)
        params = rnd.choice([(1, 3, 3, 2, 10), (1, 3, 3, 2, 10), (1, 2, 2, 2, 10), (2, 3, 4, 3, 8), (1, 1, 1, 1, 5)])
        out.append((c, r1, r2, params))
    return out


if __name__ == "__main__":
    vec = []
    for c, r1, r2, prm in cases():
        vec.append({"contig": c, "region1": r1, "region2": r2, "params": list(prm),
                    "expected": norm(ref.sw_jump(c, r1, r2, *prm))})
    with open(os.path.join(HERE, "sw_jump_vectors.json"), "w") as h:
        json.dump({"_comment": "produced by the reference's own nanomonsv/smith_waterman.sw_jump (v0.8.0); params = match_score, "
                               "mismatch_penalty, gap_cost, jump_cost, minimum_score", "vectors": vec}, h, indent=0)
    print(len(vec), "vectors,", sum(v["expected"] is None for v in vec), "None")
