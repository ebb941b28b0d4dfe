"""GPU parity tests of the sw_jump row (reference nanomonsv/smith_waterman.py:5-127): the CUDA path through
nmsv_sw_jump_batch against (a) vectors produced by the reference's own function, (b) the C oracle on larger seeded
problems of the shapes locate_bp.get_refined_bp creates (locate_bp.py:21-96)."""
import json
import os
import random

import pytest

from nanomonsv_b200 import smith_waterman as SW
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_vectors(engine):
    vec = json.load(open(os.path.join(GOLDEN, "sw_jump_vectors.json")))["vectors"]
    by_params = {}
    for v in vec:
        by_params.setdefault(tuple(v["params"]), []).append(v)
    n_found = 0
    for prm, vs in by_params.items():
        got = SW.sw_jump_batch([(v["contig"], v["region1"], v["region2"]) for v in vs], *prm, engine=engine)
        for v, g in zip(vs, got):
            if v["expected"] is None:
                assert g is None
            else:
                assert [g[0], list(g[1]), list(g[2]), list(g[3]), g[4], g[5]] == v["expected"]
                n_found += 1
    assert n_found > 80
    one = vec[0]
    g = SW.sw_jump(one["contig"], one["region1"], one["region2"], *one["params"], engine=engine)
    assert [g[0], list(g[1]), list(g[2]), list(g[3]), g[4], g[5]] == one["expected"]


def _junction_problem(rnd, n_flank, m1, m2, ins, err):
    alpha = "ACGT"
    r1 = "".join(rnd.choice(alpha) for _ in range(m1))
    r2 = "".join(rnd.choice(alpha) for _ in range(m2))
    core = r1[rnd.randint(0, 8):m1 - rnd.randint(0, 8)] + "".join(rnd.choice(alpha) for _ in range(ins)) + \
        r2[rnd.randint(0, 8):m2 - rnd.randint(0, 8)]
    c = list("".join(rnd.choice(alpha) for _ in range(rnd.randint(0, n_flank))) + core +
             "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, n_flank))))
    for _ in range(int(len(c) * err)):
        p = rnd.randrange(len(c))
        op = rnd.random()
        if op < 0.4:
            c[p] = rnd.choice(alpha)
        elif op < 0.7:
            del c[p]
        else:
            c.insert(p, rnd.choice(alpha))
    return "".join(c), r1, r2


def test_locate_bp_shaped_problems_against_c_oracle(engine):
    """Consensus contigs of a few hundred to ~2000 bases against +-20 bp breakpoint regions (mode d/r/t) and reference
    windows against 500-base contig ends (mode i), default get parameters 1/3/3/2 (arg_parser.py:121)."""
    rnd = random.Random(2026)
    probs = []
    for _ in range(40):
        probs.append(_junction_problem(rnd, 300, rnd.randint(40, 400), rnd.randint(40, 400), rnd.choice([0, 0, 5, 60]), 0.03))
    for _ in range(6):      # larger: 2 kb contig, ~1 kb regions
        probs.append(_junction_problem(rnd, 600, rnd.randint(700, 1100), rnd.randint(700, 1100), rnd.choice([0, 30]), 0.04))
    for _ in range(10):     # mode "i": the reference window plays the contig role, contig ends play the regions
        probs.append(_junction_problem(rnd, 100, 500, 500, 0, 0.02))
    for _ in range(8):      # unrelated -> None
        probs.append(("".join(rnd.choice("ACGT") for _ in range(rnd.randint(50, 500))),
                      "".join(rnd.choice("ACGT") for _ in range(rnd.randint(30, 200))),
                      "".join(rnd.choice("ACGT") for _ in range(rnd.randint(30, 200)))))
    for prm in ((1, 3, 3, 2, 10), (1, 2, 2, 2, 10)):
        got = SW.sw_jump_batch(probs, *prm, engine=engine)
        n_none = 0
        for (c, r1, r2), g in zip(probs, got):
            exp = O.sw_jump(c, r1, r2, *prm)
            assert g == exp, (len(c), len(r1), len(r2), prm)
            n_none += exp is None
        assert 4 <= n_none < len(probs) // 2


def test_edge_shapes(engine):
    cases = [("A", "A", "A"), ("ACGT", "A", "T"), ("ACGTACGTACGTACGTACGTAAAACCCCGGGGTTTTACGT", "ACGTACGTACGTACGTACGT", "AAAACCCCGGGGTTTTACGT"),
             ("N" * 30 + "ACGTTGCA" * 5, "N" * 30, "ACGTTGCA" * 5), ("acgt" * 10 + "ACGT" * 10, "acgt" * 10, "ACGT" * 10)]
    for prm in ((1, 2, 2, 2, 10), (1, 1, 1, 1, 3)):
        got = SW.sw_jump_batch(cases, *prm, engine=engine)
        for (c, r1, r2), g in zip(cases, got):
            assert g == O.sw_jump(c, r1, r2, *prm)
    assert SW.sw_jump_batch([], engine=engine) == []
    with pytest.raises(ValueError):
        SW.sw_jump("", "A", "A", engine=engine)


def test_every_strip_layout(engine):
    """The warp kernel gives lanes 0..a-1 ceil(m/32) columns and the others one fewer, walks the strip in groups of four cells and
    drops the phantom cells of the last group: region lengths around every boundary of that layout (multiples of 32 and of
    128, one more, one less, shorter than a warp), both matrices independently, against the C oracle. Lengths above 1024 take
    the block kernel; so does a scoring whose values would not fit the warp kernel's scaled 32-bit lanes."""
    rnd = random.Random(77)
    lens = [1, 2, 3, 5, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 127, 128, 129, 130, 159, 160, 161, 255, 256, 257, 383, 384, 385,
            511, 512, 513, 640, 641, 767, 768, 769, 895, 896, 897, 991, 992, 993, 1022, 1023, 1024, 1025, 1100]
    probs = []
    for k, m1 in enumerate(lens):
        m2 = lens[(k * 7 + 3) % len(lens)]
        probs.append(_junction_problem(rnd, 40, m1, m2, rnd.choice([0, 4]), 0.03))
    for m in (1, 7, 32, 33, 100):           # contig shorter than a warp: the wavefront never fills
        r1 = "".join(rnd.choice("ACGT") for _ in range(m))
        r2 = "".join(rnd.choice("ACGT") for _ in range(m + 3))
        probs.append((r1[:5] + r2[:6], r1, r2))
    for prm in ((1, 3, 3, 2, 10), (2, 1, 4, 3, 6), (1, 1, 0, 1, 3)):
        got = SW.sw_jump_batch(probs, *prm, engine=engine)
        for (c, r1, r2), g in zip(probs, got):
            assert g == O.sw_jump(c, r1, r2, *prm), (len(c), len(r1), len(r2), prm)
    big = (300000, 3, 3, 2, 10)             # match score beyond the warp kernel's range: block kernel, same answers
    few = probs[10:16]
    assert SW.sw_jump_batch(few, *big, engine=engine) == [O.sw_jump(c, r1, r2, *big) for c, r1, r2 in few]
