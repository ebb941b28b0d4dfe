"""GPU parity tests of the operator seam (parasail.ssw) and function seam (ssw_check): CUDA vs the CPU oracle and
the golden vectors. Bar: bit-exact scores and coordinates. All calls go through the C ABI (ctypes)."""
import json
import os
import random

import numpy as np
import pytest

from nanomonsv_b200 import SeqPack, ssw as gpu_ssw
from nanomonsv_b200 import _lib
from nanomonsv_b200.my_seq import needs_explicit_revcomp, reverse_complement
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _gpu_pairs(engine, pairs, scoring=(2, -2, 3, 1)):
    pack = SeqPack(dedup=True)
    t = [pack.add(a) for a, _ in pairs]
    r = [pack.add(b) for _, b in pairs]
    return engine.ssw_batch(pack, t, r, scoring)


def _assert_same(res, a, b, scoring=(2, -2, 3, 1), ctx=""):
    exp = O.ssw(a, b, scoring[2], scoring[3], scoring[0], scoring[1])
    got = (int(res["score1"]), int(res["read_begin1"]), int(res["read_end1"]), int(res["ref_begin1"]), int(res["ref_end1"]))
    assert got == exp.astuple(), (ctx, len(a), len(b), got, exp.astuple())


def test_seed_vectors(engine):
    vec = json.load(open(os.path.join(GOLDEN, "ssw_seed_vectors.json")))["vectors"]
    pairs = []
    for v in vec:
        pairs += [(v["template"], v["read"]), (O.reverse_complement(v["template"]), v["read"])]
    res = _gpu_pairs(engine, pairs)
    for k, v in enumerate(vec):
        for j, key in enumerate(("fwd", "rc")):
            r = res[2 * k + j]
            exp = v[key]
            assert int(r["score1"]) == exp[0]
            if exp[0] > 0:
                assert [int(r["score1"]), int(r["read_begin1"]), int(r["read_end1"]), int(r["ref_begin1"]), int(r["ref_end1"])] == exp
    # function seam: strand rule, tie -> '-', mirrored coordinates
    pack = SeqPack()
    t = [pack.add(v["template"]) for v in vec]
    r = [pack.add(v["read"]) for v in vec]
    chk = engine.ssw_check_batch(pack, t, r)
    for v, a in zip(vec, chk):
        got = [int(a["score"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"]), "+" if a["strand"] > 0 else "-"]
        assert got == v["check"], (v, got)


def test_fuzz_vectors_all_scorings(engine):
    vec = json.load(open(os.path.join(GOLDEN, "ssw_fuzz_vectors.json")))["vectors"]
    by_sc = {}
    for v in vec:
        by_sc.setdefault((v["match"], v["mismatch"], v["open"], v["extend"]), []).append(v)
    for sc, vs in by_sc.items():
        res = _gpu_pairs(engine, [(v["s1"], v["s2"]) for v in vs], sc)
        for v, r in zip(vs, res):
            got = [int(r["score1"]), int(r["read_begin1"]), int(r["read_end1"]), int(r["ref_begin1"]), int(r["ref_end1"])]
            if v["expected"][0] == 0:
                assert got[0] == 0
            else:
                assert got == v["expected"], (v, got)


def _mutated_copy(rnd, a, rate=0.08, alpha="ACGT"):
    b = list(a)
    for _ in range(int(len(a) * rate)):
        p = rnd.randrange(len(b))
        op = rnd.random()
        if op < 0.4:
            b[p] = rnd.choice(alpha)
        elif op < 0.7:
            del b[p]
        else:
            b.insert(p, rnd.choice(alpha))
    return "".join(b)


def test_random_ont_like_pairs_single_tile(engine):
    rnd = random.Random(101)
    pairs = []
    for _ in range(120):
        m = rnd.choice([400, 400, 399, 401, 416, 417, 200, 50, 13, 1, 601])
        a = "".join(rnd.choice("ACGT") for _ in range(m))
        kind = rnd.random()
        if kind < 0.6:
            core = _mutated_copy(rnd, a if rnd.random() < 0.5 else O.reverse_complement(a))
            b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 3000))) + core + \
                "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 3000)))
        elif kind < 0.8:
            b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 4000)))
        else:   # partial overlap at the read end / read shorter than template
            cut = rnd.randint(1, m)
            b = _mutated_copy(rnd, a[:cut]) or "A"
        pairs.append((a, b))
    res = _gpu_pairs(engine, pairs)
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b)


def test_tie_heavy_and_wildcards(engine):
    rnd = random.Random(202)
    pairs = []
    for alpha in ("A", "AC", "AAAC", "ACGTN", "ACN"):
        for _ in range(40):
            a = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 450)))
            b = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 900)))
            pairs.append((a, b))
    pairs += [("N" * 50, "ACGT" * 20), ("ACGT" * 20, "N" * 70), ("A" * 400, "A" * 1000), ("ACGT" * 100, "ACGT" * 300),
              ("AC" * 200, "CA" * 333), ("A", "A"), ("A", "C"), ("acgtACGT" * 20, "ACGTacgt" * 41)]
    res = _gpu_pairs(engine, pairs)
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b)


@pytest.mark.parametrize("m", [1000, 2000, 2048, 2049, 6000])
def test_multi_tile_templates(engine, m):
    rnd = random.Random(300 + m)
    pairs = []
    for k in range(6):
        a = "".join(rnd.choice("ACGT") for _ in range(m))
        src = a if k % 2 == 0 else O.reverse_complement(a)
        if k == 4:
            src = src[: m // 2]                      # only half of the template is in the read
        core = _mutated_copy(rnd, src)
        b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 2500))) + core + \
            "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 2500)))
        if k == 5:
            b = "".join(rnd.choice("ACGT") for _ in range(3000))   # unrelated read
        pairs.append((a, b))
    res = _gpu_pairs(engine, pairs)
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b, ctx=f"m={m}")


def test_long_reads_up_to_50kb(engine):
    rnd = random.Random(404)
    pairs = []
    for n in (20000, 50000):
        a = "".join(rnd.choice("ACGT") for _ in range(400))
        body = [rnd.choice("ACGT") for _ in range(n)]
        at = n - 5000
        core = _mutated_copy(rnd, O.reverse_complement(a))
        b = "".join(body[:at]) + core + "".join(body[at:])
        pairs.append((a, b))
        pairs.append((a, "".join(body)))
    res = _gpu_pairs(engine, pairs)
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b)


def test_begin_window_covers_long_insertions(engine):
    """The begin pass visits a bounded window of reversed columns (rows + (match*rows - score) / min(open, ext)): reads
    that carry a long insertion inside the template -- the optimal local alignment spans far more columns than rows --
    must still give the begin cell of the full rectangle."""
    rnd = random.Random(909)
    pairs, tmpls = [], []
    for m, ins in ((600, 250), (400, 150), (2000, 700), (1200, 160), (1200, 140)):
        a = "".join(rnd.choice("ACGT") for _ in range(m))
        cut = m // 2 + rnd.randint(-20, 20)
        for strand in (1, -1):
            src = a if strand > 0 else O.reverse_complement(a)
            b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(100, 900))) + _mutated_copy(rnd, src[:cut], 0.03) + \
                "".join(rnd.choice("ACGT") for _ in range(ins)) + _mutated_copy(rnd, src[cut:], 0.03) + \
                "".join(rnd.choice("ACGT") for _ in range(rnd.randint(100, 900)))
            pairs.append((src, b))
            tmpls.append(a)
    res = _gpu_pairs(engine, pairs)
    spans = 0
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b)
        rows = int(r["read_end1"]) - int(r["read_begin1"]) + 1
        cols = int(r["ref_end1"]) - int(r["ref_begin1"]) + 1
        spans += cols > rows + rows // 8 + 64
    assert spans >= 6          # the insertion really is inside most alignments
    pack = SeqPack()
    t = [pack.add(a) for a in tmpls]          # the template itself: every second read holds its reverse complement
    rd = [pack.add(b) for _, b in pairs]
    chk = engine.ssw_check_batch(pack, t, rd)
    for a, (_, b), c in zip(tmpls, pairs, chk):
        exp = O.ssw_check(a, [("r", b)])["r"]
        assert [int(c["score"]), int(c["qstart"]), int(c["qend"]), int(c["tstart"]), int(c["tend"]), "+" if c["strand"] > 0 else "-"] == exp
    assert sorted(int(c["strand"]) for c in chk) == [-1] * 5 + [1] * 5


@pytest.mark.parametrize("scoring", [(1, -2, 3, 1), (2, -3, 5, 2), (3, -1, 2, 2), (2, -2, 3, 3), (2, 0, 4, 1),
                                     (2, -2, 4, 0), (1, -1, 0, 1), (2, -2, 0, 0)])
def test_other_scorings(engine, scoring):
    """The other parasail.ssw call sites use other matrices (locate_bp_sbnd.py:28 uses 1/-2)."""
    rnd = random.Random(hash(scoring) & 0xffff)
    pairs = []
    for _ in range(40):
        a = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(20, 500)))
        b = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 300))) + _mutated_copy(rnd, a, 0.12) + \
            "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 300)))
        pairs.append((a, b))
    res = _gpu_pairs(engine, pairs, scoring)
    for (a, b), r in zip(pairs, res):
        _assert_same(r, a, b, scoring)


def test_ssw_check_matches_oracle_including_lowercase_templates(engine):
    rnd = random.Random(505)
    templates, reads = [], []
    for k in range(30):
        alpha = "ACGT" if k % 3 else "ACGTacgtN"
        t = "".join(rnd.choice(alpha) for _ in range(rnd.randint(30, 420)))
        src = t.upper() if rnd.random() < 0.5 else O.reverse_complement(t.upper())
        rd = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 800))) + _mutated_copy(rnd, src) + \
             "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 800)))
        templates.append(t)
        reads.append(rd)
    pack = SeqPack()
    ti = [pack.add(t) for t in templates]
    rci = [pack.add(reverse_complement(t)) if needs_explicit_revcomp(t) else _lib.NMSV_NONE for t in templates]
    ri = [pack.add(r) for r in reads]
    res = engine.ssw_check_batch(pack, ti, ri, rci)
    for t, rd, a in zip(templates, reads, res):
        exp = O.ssw_check(t, [("r", rd)])["r"]
        got = [int(a["score"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"]), "+" if a["strand"] > 0 else "-"]
        assert got == exp, (t[:30], got, exp)


def test_parasail_shaped_facade(engine):
    m = gpu_ssw.matrix_create("ACGT", 2, -2)
    r = gpu_ssw.ssw("ACGTACGTAC", "ACGTCGTAC", 3, 1, m, engine=engine)
    assert (r.score1, r.read_begin1, r.read_end1, r.ref_begin1, r.ref_end1) == (15, 0, 9, 0, 8)
    # 16-bit range: parasail returns None; so does the facade
    assert gpu_ssw.ssw("A" * 16400, "A" * 16400, 3, 1, m, engine=engine) is None
    with pytest.raises(ValueError):
        gpu_ssw.matrix_create("ARND", 2, -2)


def test_errors_are_loud(engine):
    from nanomonsv_b200 import EngineError
    pack = SeqPack()
    a = pack.add("ACGT")
    with pytest.raises(EngineError):
        engine.ssw_batch(pack, [a], [7])                 # index out of range
    with pytest.raises(EngineError) as e:
        engine.ssw_batch(pack, [a], [a], (100, -2, 3, 1))   # scoring parameters beyond the 16-bit lanes
    assert e.value.code == _lib.E_RANGE
    # a long template is no error: only a score that leaves the 16-bit range is, and then per pair (None)
    pack2 = SeqPack()
    big = pack2.add_codes(np.zeros(17000, dtype=np.uint8))
    rd = pack2.add("ACGT")
    assert int(engine.ssw_batch(pack2, [big], [rd])[0]["score1"]) == 2


def test_generate_paf_file_shape(engine):
    """generate_consensus.generate_paf_file (reference generate_consensus.py:100-127) calls parasail.ssw(qseq, tseq, 3, 1,
    M) with s1 = a WHOLE READ (10-50 kb) and s2 = the consensus (400-2000 b), and handles None per pair (:120-125).
    The score is bounded by 2 * len(s2): no such pair may come back None, whatever len(s1) -- including the lengths
    16 293 .. 16 380 that the round-1 length rule treated inconsistently."""
    rnd = random.Random(91)
    m = gpu_ssw.matrix_create("ACGT", 2, -2)
    pairs = []
    for n1 in (10000, 16293, 16300, 16380, 16381, 23000, 50000):
        cons = "".join(rnd.choice("ACGT") for _ in range(rnd.choice((400, 900, 2000))))
        body = _mutated_copy(rnd, cons, 0.08)
        at = rnd.randrange(0, n1 - len(body))
        read = "".join(rnd.choice("ACGT") for _ in range(at)) + body
        read += "".join(rnd.choice("ACGT") for _ in range(n1 - len(read)))
        pairs.append((read, cons))
        pairs.append((O.reverse_complement(read), cons))        # unrelated strand: a low score, still a result
    got = gpu_ssw.ssw_batch(pairs, 3, 1, m, engine=engine)
    for (s1, s2), g in zip(pairs, got):
        e = O.ssw(s1, s2, engine="striped")
        assert g is not None and e is not None
        assert (g.score1, g.read_begin1, g.read_end1, g.ref_begin1, g.ref_end1) == e.astuple(), (len(s1), len(s2))
    assert sum(1 for g in got if g.score1 > 500) == len(pairs) // 2
    # long against long: a result while the score is representable, None once it is not (like the oracle / parasail)
    a = "".join(rnd.choice("ACGT") for _ in range(17000))
    b = "".join(rnd.choice("ACGT") for _ in range(19000))
    g = gpu_ssw.ssw(a, b, 3, 1, m, engine=engine)
    assert g is not None and (g.score1, g.read_begin1, g.read_end1, g.ref_begin1, g.ref_end1) == O.ssw(a, b, engine="striped").astuple()
    assert gpu_ssw.ssw(a, a, 3, 1, m, engine=engine) is None and O.ssw(a, a, engine="striped") is None
