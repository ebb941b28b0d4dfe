"""GPU tests at the FULL sizes of the BASELINE configs (templates of 2 kb and 6 kb, reads up to 50 kb, 601-base
single-breakend templates) where the CPU oracle would take too long for every pair: size-independent properties of
the alignment, plus oracle spot checks on a few pairs per shape."""

import numpy as np
import pytest

from nanomonsv_b200 import SeqPack, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rand(rng, n):
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def _noisy(rng, codes, rate=0.08):
    cfg = synth.SynthConfig(sub_rate=rate * 3 / 8, ins_rate=rate * 2.5 / 8, del_rate=rate * 2.5 / 8)
    out, st, ln = synth.mutate_concat(rng, codes, np.zeros(1, dtype=np.int64), cfg)
    return out[: int(ln[0])]


@pytest.mark.parametrize("m,n_lo,n_hi", [(2000, 8000, 50000), (6000, 8000, 20000), (601, 5000, 50000), (400, 1000, 50000)])
def test_full_size_properties(engine, m, n_lo, n_hi):
    rng = np.random.default_rng(m * 7 + 1)
    pack = SeqPack()
    t_idx, r_idx, meta = [], [], []
    n_pairs = 24
    for i in range(n_pairs):
        tmpl = _rand(rng, m)
        n = int(rng.integers(n_lo, n_hi + 1))
        body = _rand(rng, n)
        kind = i % 4
        at = int(rng.integers(0, n - m))
        if kind == 0:       # exact forward copy of the template inside the read
            body[at:at + m] = tmpl
        elif kind == 1:     # exact reverse-complement copy
            body[at:at + m] = synth.revcomp_codes(tmpl)
        elif kind == 2:     # noisy copy (8 % error) -> a high but imperfect score
            noisy = _noisy(rng, tmpl)
            body = np.concatenate((body[:at], noisy, body[at:]))
        ti, ri = pack.add_codes(tmpl), pack.add_codes(body)
        t_idx.append(ti); r_idx.append(ri); meta.append((kind, at, tmpl, body))
    chk = engine.ssw_check_batch(pack, t_idx, r_idx)
    fwd = engine.ssw_batch(pack, t_idx, r_idx)
    chk2 = engine.ssw_check_batch(pack, t_idx, r_idx)
    assert (chk == chk2).all()                                   # idempotent / deterministic
    for (kind, at, tmpl, body), a, f in zip(meta, chk, fwd):
        n = len(body)
        assert 0 <= a["score"] <= 2 * m
        assert 1 <= a["qstart"] <= a["qend"] <= m and 1 <= a["tstart"] <= a["tend"] <= n
        if kind == 0:
            assert (a["score"], a["strand"], a["qstart"], a["qend"]) == (2 * m, 1, 1, m)
            assert a["tend"] - a["tstart"] == m - 1 and a["tstart"] <= at + 1   # first occurrence wins
            assert f["score1"] == 2 * m and f["read_end1"] == m - 1
        elif kind == 1:
            assert (a["score"], a["strand"], a["qstart"], a["qend"]) == (2 * m, -1, 1, m)
        elif kind == 2:
            assert a["score"] > 1.2 * m and a["strand"] == 1 and a["qstart"] < 0.1 * m and a["qend"] > 0.9 * m
            assert f["score1"] == a["score"]                      # the forward strand is the winner
        else:
            # unrelated sequences: with +2/-2 and gaps 3/1 random DNA is in the linear regime of local alignment
            # (about 0.2 * m), still far below the 1.2 * m support threshold of the reference
            assert a["score"] < 0.45 * m
    # oracle spot checks on the two smallest reads of every kind
    for kind in range(4):
        idx = sorted((i for i in range(n_pairs) if meta[i][0] == kind), key=lambda i: len(meta[i][3]))[:2]
        for i in idx:
            t, b = synth.codes_to_str(meta[i][2]), synth.codes_to_str(meta[i][3])
            exp = O.ssw_check(t, [("r", b)], engine="striped")["r"]
            a = chk[i]
            assert [int(a["score"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"]),
                    "+" if a["strand"] > 0 else "-"] == exp


def test_strand_symmetry_and_truncation_monotonicity(engine):
    """ssw_check(T, R) and ssw_check(revcomp(T), R) see the same two alignments with swapped roles; cutting the read
    can never raise a score; appending unrelated bases can never lower it."""
    rng = np.random.default_rng(99)
    pack = SeqPack()
    rows = []
    for _ in range(20):
        m = int(rng.integers(300, 2100))
        tmpl = _rand(rng, m)
        n = int(rng.integers(3000, 20000))
        body = _rand(rng, n)
        at = int(rng.integers(0, n - m))
        body = np.concatenate((body[:at], _noisy(rng, tmpl if rng.random() < 0.5 else synth.revcomp_codes(tmpl)), body[at:]))
        t = pack.add_codes(tmpl)
        trc = pack.add_codes(synth.revcomp_codes(tmpl))
        r_full = pack.add_codes(body)
        r_cut = pack.add_codes(body[: len(body) // 2])
        r_ext = pack.add_codes(np.concatenate((body, _rand(rng, 500))))
        rows.append((t, trc, r_full, r_cut, r_ext))
    res = engine.ssw_batch(pack, [x for row in rows for x in (row[0], row[1], row[0], row[0])],
                           [x for row in rows for x in (row[2], row[2], row[3], row[4])])
    chk = engine.ssw_check_batch(pack, [row[0] for row in rows] + [row[1] for row in rows],
                                 [row[2] for row in rows] * 2)
    k = len(rows)
    for i in range(k):
        f_t, f_rc, f_cut, f_ext = res[4 * i: 4 * i + 4]
        a, b = chk[i], chk[k + i]
        assert max(f_t["score1"], f_rc["score1"]) == a["score"] == b["score"]
        if f_t["score1"] != f_rc["score1"]:
            assert a["strand"] == -b["strand"]
            assert (a["tstart"], a["tend"]) == (b["tstart"], b["tend"])
        assert f_cut["score1"] <= f_t["score1"] <= f_ext["score1"]


def test_config5_shaped_sweep_slice(engine):
    """A slice of the kernel sweep of BASELINE config 5: m = 400, read lengths spread over 1-50 kb, pairs drawn from a
    pool; the pair API, the function seam and the oracle agree on a sample."""
    seqs, t_idx, r_idx = synth.make_pair_sweep(600, m=400, n_min=1000, n_max=50000, pool_reads=120, seed=11)
    pack = SeqPack()
    for s in seqs:
        pack.add_codes(s)
    res = engine.ssw_batch(pack, t_idx, r_idx)
    chk = engine.ssw_check_batch(pack, t_idx, r_idx)
    assert (chk["score"] >= res["score1"]).all()
    assert ((chk["strand"] == 1) == (chk["score"] == res["score1"]))[chk["strand"] == 1].all()
    hits = int((chk["score"] > 480).sum())
    assert 0.3 * len(t_idx) < hits < 0.7 * len(t_idx)            # about half of the reads carry the template
    order = np.argsort([len(seqs[r]) for r in r_idx])[:12]
    for p in order:
        t, b = synth.codes_to_str(seqs[t_idx[p]]), synth.codes_to_str(seqs[r_idx[p]])
        e = O.ssw(t, b, engine="striped")
        got = (int(res[p]["score1"]), int(res[p]["read_begin1"]), int(res[p]["read_end1"]), int(res[p]["ref_begin1"]), int(res[p]["ref_end1"]))
        assert got == e.astuple()


def test_bias_rebase_on_very_long_reads(engine):
    """Reads long enough to force several rebases of the step bias inside the kernel (m = 6000 leaves ~20 k steps of
    16-bit head-room): results must not depend on where the rebases fall."""
    rng = np.random.default_rng(5)
    tmpl = _rand(rng, 6000)
    body = _rand(rng, 70000)
    body[61000:67000] = tmpl
    pack = SeqPack()
    t, r = pack.add_codes(tmpl), pack.add_codes(body)
    a = engine.ssw_check_batch(pack, [t], [r])[0]
    assert (int(a["score"]), int(a["strand"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"])) == \
           (12000, 1, 1, 6000, 61001, 67000)
    tmpl2 = _rand(rng, 14000)             # near the 16-bit limit: 2 * 14000 = 28000 -> a rebase every ~4.7 k steps
    body2 = np.concatenate((_rand(rng, 9000), tmpl2, _rand(rng, 9000)))
    pack = SeqPack()
    t, r = pack.add_codes(tmpl2), pack.add_codes(body2)
    f = engine.ssw_batch(pack, [t], [r])[0]
    assert (int(f["score1"]), int(f["read_begin1"]), int(f["read_end1"]), int(f["ref_begin1"]), int(f["ref_end1"])) == \
           (28000, 0, 13999, 9000, 22999)


def test_no_length_limit_only_scores_saturate(engine):
    """parasail.ssw returns None only when a SCORE leaves the 16-bit range, whatever the lengths. The kernel runs any
    task whose bound match * min(rows, columns) fits below sw_score_cap unchecked (the longest such template forces a
    bias rebase at every chunk and must be exact) and watches longer ones at run time."""
    rng = np.random.default_rng(17)
    # sw_score_cap = 32767 - match - 1 - headroom, headroom (nmsv_sw.cuh: sw_headroom) = bias (open + 7 * ext + 1)
    # + ext + (128 + 8 + 32) * ext = 180
    lmax = (32767 - 2 - 1 - 180) // 2
    tmpl = _rand(rng, lmax)
    body = np.concatenate((_rand(rng, 700), tmpl, _rand(rng, 900)))
    pack = SeqPack()
    t, r, trc = pack.add_codes(tmpl), pack.add_codes(body), pack.add_codes(synth.revcomp_codes(tmpl))
    f = engine.ssw_batch(pack, [t], [r])[0]
    assert (int(f["score1"]), int(f["read_begin1"]), int(f["read_end1"]), int(f["ref_begin1"]), int(f["ref_end1"])) == \
           (2 * lmax, 0, lmax - 1, 700, 700 + lmax - 1)
    a = engine.ssw_check_batch(pack, [trc], [r])[0]
    assert (int(a["score"]), int(a["strand"]), int(a["qstart"]), int(a["qend"])) == (2 * lmax, -1, 1, lmax)
    # longer templates: exact as long as the score stays below the cap ...
    long_t = _rand(rng, 20000)
    part = np.concatenate((_rand(rng, 300), long_t[4000:9000], _rand(rng, 500)))
    unrelated = _rand(rng, 18000)
    # ... and None (score1 = -1) when it does not: a 16 400-base perfect match scores 32 800 > 32 764
    perfect = _rand(rng, 16400)
    pack = SeqPack()
    ti, pi, ui, qi = pack.add_codes(long_t), pack.add_codes(part), pack.add_codes(unrelated), pack.add_codes(perfect)
    res = engine.ssw_batch(pack, [ti, ti, qi, pi], [pi, ui, qi, ti])
    codes, offsets, lens = pack.finalize()
    exp, _ = O.ssw_batch(codes, offsets, lens, np.array([ti, ti, qi, pi]), np.array([pi, ui, qi, ti]), engine="striped")
    for k in (0, 1, 3):
        assert tuple(int(res[k][n]) for n in res.dtype.names) == tuple(int(exp[k][n]) for n in exp.dtype.names), k
    assert int(res[0]["score1"]) > 8000
    assert int(exp[2]["score1"]) == -1 and tuple(int(res[2][n]) for n in res.dtype.names) == (-1,) * 5

