"""CPU tests of the oracle (test infrastructure) against the golden vectors: -m "not gpu"."""
import json
import os
import random

import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _seed():
    return json.load(open(os.path.join(GOLDEN, "ssw_seed_vectors.json")))["vectors"]


def _fuzz():
    return json.load(open(os.path.join(GOLDEN, "ssw_fuzz_vectors.json")))["vectors"]


@pytest.mark.parametrize("engine", ["scalar", "striped", "striped16"])
def test_seed_vectors(engine):
    for v in _seed():
        for tmpl, key in ((v["template"], "fwd"), (O.reverse_complement(v["template"]), "rc")):
            got = O.ssw(tmpl, v["read"], engine=engine).astuple()
            exp = v[key]
            assert got[0] == exp[0], (v, key, got)
            if exp[0] > 0:
                assert list(got) == exp, (v, key, got)
        info = O.ssw_check(v["template"], [("r", v["read"])], engine=engine)["r"]
        assert info == v["check"], (v, info)


@pytest.mark.parametrize("engine", ["scalar", "striped", "striped16"])
def test_fuzz_vectors_from_pure_python(engine):
    for v in _fuzz():
        if engine != "scalar" and v["open"] <= v["extend"]:
            # Farrar's lazy-F early exit (kept in the striped port because parasail has it) is only exact when a
            # gap's first base costs more than an extension; the validation path always uses open=3 > extend=1.
            continue
        got = O.ssw(v["s1"], v["s2"], v["open"], v["extend"], v["match"], v["mismatch"], engine=engine).astuple()
        if v["expected"][0] == 0:
            assert got[0] == 0
        else:
            assert list(got) == v["expected"], (v, got)


def test_scalar_equals_striped_on_long_high_scoring_pairs():
    rnd = random.Random(11)
    saturating = 0
    for _ in range(40):
        m = rnd.randint(150, 900)
        a = "".join(rnd.choice("ACGT") for _ in range(m))
        b = list(a)
        for _ in range(int(m * 0.08)):
            p = rnd.randrange(len(b))
            op = rnd.random()
            if op < 0.4:
                b[p] = rnd.choice("ACGT")
            elif op < 0.7:
                del b[p]
            else:
                b.insert(p, rnd.choice("ACGT"))
        b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 700))) + "".join(b) + \
            "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 700)))
        if rnd.random() < 0.5:
            a = O.reverse_complement(a)
        r = [O.ssw(a, b, engine=e).astuple() for e in ("scalar", "striped", "striped16")]
        assert r[0] == r[1] == r[2]
        saturating += r[0][0] > 252   # exercised the 8-bit saturation -> 16-bit rerun
    assert saturating > 20


def test_saturation_returns_none():
    a = "A" * 16400
    assert O.ssw(a, a) is None
    assert O.ssw(a, a, engine="striped") is None
    assert O.ssw("A" * 100, "A" * 100) is not None


def test_wildcard_and_case():
    assert O.ssw("ACGTNACGT", "ACGTAACGT").score1 == 16
    assert O.ssw("acgtnacgt", "ACGTAACGT").astuple() == O.ssw("ACGTNACGT", "ACGTAACGT").astuple()
    assert O.ssw("NNNN", "ACGT").score1 == 0


def test_reverse_complement_quirks():
    assert O.reverse_complement("ACGTN") == "NACGT"
    assert O.reverse_complement("AcgT") == "AgcT"       # lower case is reversed but not complemented
    assert O.reverse_complement("MRWX") == "XWYK"


def test_batch_matches_single():
    import numpy as np
    seqs = ["ACGTACGTAC", "ACGTCGTAC", "GATTACA", "TGTAATC"]
    codes = np.concatenate([O.encode(s) for s in seqs])
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offs = np.concatenate(([0], np.cumsum(lens[:-1]))).astype(np.uint64)
    for eng in ("scalar", "striped"):
        res, used = O.ssw_batch(codes, offs, lens, [0, 2], [1, 3], engine=eng, threads=2)
        assert tuple(res[0]) == (15, 0, 8, 0, 9) and tuple(res[1]) == (4, 2, 3, 3, 4)
        assert used >= 1


def test_validate_fixture_is_reproduced_by_the_oracle(tmp_path):
    fix = json.load(open(os.path.join(GOLDEN, "validate_small.json")))
    fasta = O.DictFasta(fix["contigs"])
    for sample, d in fix["samples"].items():
        seam = tmp_path / (sample + ".tsv")
        seam.write_text("\n".join(d["seam"]) + "\n")
        O.count_from_seam(str(seam), fasta, str(seam) + ".count", str(seam) + ".info", var_read_min_mapq=40,
                          engine="striped")
        assert open(str(seam) + ".count").read().splitlines() == d["count"]
        assert sorted(open(str(seam) + ".info").read().splitlines()) == d["info"]


def test_sw_jump_oracle_equals_reference_vectors():
    """The vectors were produced by the reference's own smith_waterman.sw_jump (tests/golden/make_golden_jump.py)."""
    vec = json.load(open(os.path.join(GOLDEN, "sw_jump_vectors.json")))["vectors"]
    assert len(vec) > 100 and any(v["expected"] is None for v in vec)
    for v in vec:
        got = O.sw_jump(v["contig"], v["region1"], v["region2"], *v["params"])
        exp = v["expected"]
        if exp is None:
            assert got is None
        else:
            assert [got[0], list(got[1]), list(got[2]), list(got[3]), got[4], got[5]] == exp
