"""Opt-in differential test against the REAL parasail (the reference's engine, pip parasail==1.3.4). parasail is not
installable in the build image (no network), so this test is skipped there; run it wherever parasail exists to pin
the oracle (and with a GPU, the CUDA path) to the real thing."""
import random

import pytest

parasail = pytest.importorskip("parasail")

from oracle import oracle as O  # noqa: E402


def _cases(n=300, seed=5):
    rnd = random.Random(seed)
    for _ in range(n):
        alpha = rnd.choice(["AC", "ACGT", "ACGTN", "AAAC"])
        a = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 500)))
        b = list(a if rnd.random() < 0.5 else O.reverse_complement(a))
        for _ in range(rnd.randint(0, max(1, len(a) // 10))):
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
        b = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 800))) + "".join(b) + \
            "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 800)))
        yield a, b or "A"


def test_oracle_equals_real_parasail():
    m = parasail.matrix_create("ACGT", 2, -2)
    for a, b in _cases():
        ref = parasail.ssw(a, b, 3, 1, m)
        got = O.ssw(a, b)
        if ref.score1 == 0:
            assert got.score1 == 0
            continue
        assert (got.score1, got.read_begin1, got.read_end1, got.ref_begin1, got.ref_end1) == \
               (ref.score1, ref.read_begin1, ref.read_end1, ref.ref_begin1, ref.ref_end1), (a, b)
