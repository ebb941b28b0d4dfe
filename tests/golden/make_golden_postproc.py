"""Regenerates tests/golden/proc_sread_info_vectors.json with the REFERENCE'S OWN proc_sread_info_file
(nanomonsv/post_proc.py:306-340). The module imports pysam and parasail at the top (both absent here), so the one
function is taken out of the source with `ast` and compiled on its own -- it only needs open/print. Nothing of the
reference's source is written into this repository: only inputs and the outputs the reference produced.

    python tests/golden/make_golden_postproc.py        (build container, /root/reference present)

The inputs are the supporting-read info lines of tests/golden/validate_small.json (the oracle's, which the CUDA path
reproduces byte for byte) plus synthetic lines for the branches those do not reach: score ties, both strands of both
variant segments, inserted sequences, lines whose key is not in the result file, a header-only result file.
"""
import ast
import json
import os
import random
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nanomonsv/post_proc.py"


def load_reference_function(name):
    tree = ast.parse(open(SRC).read(), SRC)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name]
    assert len(fn) == 1, name
    mod = ast.Module(body=fn, type_ignores=[])
    ns = {}
    exec(compile(mod, SRC, "exec"), ns)
    return ns[name]


def synthetic_lines(rnd, n):
    out = []
    for i in range(n):
        inseq = rnd.choice(["---", "---", "ACGTTGCA" * rnd.randint(1, 30), "A"])
        key = ["chr%d" % rnd.randint(1, 3), str(rnd.randint(1000, 90000)), rnd.choice("+-"), "chr%d" % rnd.randint(1, 3),
               str(rnd.randint(1000, 90000)), rnd.choice("+-"), inseq, "c_%d" % rnd.randint(0, 9)]
        L = rnd.choice([200, 200, 1000])

        def aln():
            m = 2 * L
            qs = rnd.randint(1, m // 10)
            qe = rnd.randint(m - m // 10, m)
            ts = rnd.randint(1, 20000)
            te = ts + (qe - qs) + rnd.randint(-30, 30)
            return [rnd.randint(int(1.2 * m), 2 * m), qs, qe, ts, te, rnd.choice("+-")]
        a1, a2 = aln(), aln()
        if i % 7 == 0:
            a2[0] = a1[0]                    # tie: the first segment wins (score1 >= score2)
        fields = key + ["read_%d" % i] + [str(x) for x in a1 + a2]
        if i % 3 == 0:
            fields += ["None"] * 12          # the short-del/dup columns of the info line are ignored
        out.append(("\t".join(fields), L))
    return out


def main():
    ref = load_reference_function("proc_sread_info_file")
    rnd = random.Random(20261020)
    fix = json.load(open(os.path.join(HERE, "validate_small.json")))
    cases = []
    real = fix["samples"]["tumor"]["info"]
    keys = sorted({"\t".join(l.split("\t")[:8]) for l in real})
    groups = [("validate_small: all keys called", real, keys, 200),
              ("validate_small: half of the keys called", real, keys[::2], 200),
              ("validate_small: header-only result file", real, [], 200)]
    for L in (200, 1000):
        syn = [l for l, ll in synthetic_lines(rnd, 120) if ll == L]
        skeys = sorted({"\t".join(l.split("\t")[:8]) for l in syn})
        groups.append((f"synthetic L={L}", syn, [k for i, k in enumerate(skeys) if i % 4 != 3], L))
    with tempfile.TemporaryDirectory() as tmp:
        for name, info, called, L in groups:
            fi, fr, fo = (os.path.join(tmp, x) for x in ("info.txt", "result.txt", "out.txt"))
            open(fi, "w").write("".join(l + "\n" for l in info))
            # result file: header line, then the key columns followed by further columns (read counts ...)
            open(fr, "w").write("Chr_1\tPos_1\tDir_1\tChr_2\tPos_2\tDir_2\tInserted_Seq\tSV_ID\tChecked_Read_Num_Tumor\n" +
                                "".join(k + "\t12\t5\n" for k in called))
            ref(fi, fr, fo, L)
            cases.append({"name": name, "validate_sequence_length": L, "info": info, "result_keys": called,
                          "expected": open(fo).read().splitlines()})
    json.dump({"source": "nanomonsv/post_proc.py:306-340 (proc_sread_info_file), executed from /root/reference",
               "cases": cases}, open(os.path.join(HERE, "proc_sread_info_vectors.json"), "w"), indent=0)
    print("cases", len(cases), "expected lines", sum(len(c["expected"]) for c in cases))


if __name__ == "__main__":
    main()
