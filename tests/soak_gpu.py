#!/usr/bin/env python
"""Randomised parity soak (not collected by pytest: run by hand on a GPU box).

    python tests/soak_gpu.py [--seconds 120] [--seed 1]

Draws batches of (template, read) pairs over a wide range of shapes -- 1 to 14 000 template rows (every strip class,
1 to 28 tiles), 1 to 40 000 columns, related / unrelated / low-complexity / tandem-repeat / soft-masked / N-rich
sequences, several scorings incl. gap_open <= gap_extend (generic kernel instantiation) -- and compares
nmsv_ssw_batch (score + end + begin cell) and nmsv_ssw_check_batch (strand rule, 1-based tuples) with the CPU oracle
(oracle/, striped AVX2 on all cores where it is exact, scalar otherwise). Exits non-zero on the first mismatch.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nanomonsv_b200 import Engine, SeqPack   # noqa: E402
from oracle import oracle as O               # noqa: E402

COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def mutate(rng, a, rate):
    """i.i.d. substitutions / insertions / deletions at `rate` per base."""
    if rate <= 0 or len(a) == 0:
        return a.copy()
    r = rng.random(len(a))
    keep = r >= rate / 3                               # deletions
    out = a[keep].copy()
    sub = rng.random(len(out)) < rate / 3
    out[sub] = rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)
    n_ins = int(rng.binomial(len(out) + 1, rate / 3))
    if n_ins:
        pos = np.sort(rng.integers(0, len(out) + 1, size=n_ins))
        out = np.insert(out, pos, rng.integers(0, 4, size=n_ins, dtype=np.uint8))
    return out


def draw_pair(rng):
    kind = rng.integers(0, 10)
    m = int(rng.choice([rng.integers(1, 50), rng.integers(50, 700), rng.integers(380, 620), rng.integers(1900, 2100),
                        rng.integers(700, 6200), rng.integers(6200, 14000) if rng.random() < 0.15 else rng.integers(300, 500)]))
    alpha = int(rng.choice([4, 4, 4, 2, 1, 5]))        # 5: with wildcards
    t = rng.integers(0, alpha, size=m, dtype=np.uint8)
    if kind == 0:                                       # tandem repeat template: many equal-score cells
        unit = rng.integers(0, 4, size=int(rng.integers(1, 12)), dtype=np.uint8)
        t = np.resize(unit, m)
    n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 5000), rng.integers(5000, 40000)]))
    body = rng.integers(0, alpha, size=n, dtype=np.uint8)
    if kind <= 6:                                       # the read carries (part of) the template, either strand
        src = t if rng.random() < 0.5 else COMP[t[::-1]]
        lo = int(rng.integers(0, max(1, m // 3))) if rng.random() < 0.3 else 0
        hi = m - (int(rng.integers(0, max(1, m // 3))) if rng.random() < 0.3 else 0)
        core = mutate(rng, src[lo:hi], float(rng.choice([0.0, 0.02, 0.08, 0.15, 0.3])))
        if rng.random() < 0.2 and len(core) > 10:       # a long insertion inside
            at = int(rng.integers(1, len(core)))
            core = np.concatenate((core[:at], rng.integers(0, 4, size=int(rng.integers(20, 400)), dtype=np.uint8), core[at:]))
        at = int(rng.integers(0, n + 1))
        body = np.concatenate((body[:at], core, body[at:]))[: max(1, n if rng.random() < 0.1 else 10 ** 9)]
    if rng.random() < 0.1 and len(body) > 4:
        k = int(rng.integers(1, max(2, len(body) // 20)))
        body[rng.integers(0, len(body), size=k)] = 4    # scattered N
    return t, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--batch-cells", type=float, default=1.5e11)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    eng = Engine(0)
    scorings = [(2, -2, 3, 1)] * 4 + [(1, -2, 3, 1), (2, -3, 5, 2), (3, -1, 2, 2), (2, -2, 1, 3), (2, 0, 4, 1), (1, -1, 0, 1), (2, -2, 4, 0)]
    t_end = time.time() + args.seconds
    n_pairs = n_batches = 0
    cells_total = 0.0
    while time.time() < t_end:
        sc = scorings[int(rng.integers(0, len(scorings)))]
        pack = SeqPack()
        t_idx, r_idx = [], []
        cells = 0.0
        budget = args.batch_cells if sc[2] > sc[3] else args.batch_cells / 60     # the scalar oracle is single-threaded
        while cells < budget and len(t_idx) < 4000:
            t, b = draw_pair(rng)
            if sc[0] * len(t) > 32000:
                continue
            t_idx.append(pack.add_codes(t)); r_idx.append(pack.add_codes(b))
            cells += 2.0 * len(t) * len(b)
        got = eng.ssw_batch(pack, t_idx, r_idx, sc)
        chk = eng.ssw_check_batch(pack, t_idx, r_idx, None, sc)
        codes, offsets, lens = pack.finalize()
        exact_striped = sc[2] > sc[3]
        # forward strand and reverse-complement strand through the oracle
        opack = SeqPack()
        of, orc, ordd = [], [], []
        for ti, ri in zip(t_idx, r_idx):
            tc = codes[int(offsets[ti]):int(offsets[ti]) + int(lens[ti])]
            of.append(opack.add_codes(tc)); orc.append(opack.add_codes(COMP[tc[::-1]]))
            ordd.append(opack.add_codes(codes[int(offsets[ri]):int(offsets[ri]) + int(lens[ri])]))
        oc, oo, ol = opack.finalize()
        engine = "striped" if exact_striped else "scalar"
        ef, _ = O.ssw_batch(oc, oo, ol, np.asarray(of), np.asarray(ordd), sc[2], sc[3], sc[0], sc[1], engine=engine)
        er, _ = O.ssw_batch(oc, oo, ol, np.asarray(orc), np.asarray(ordd), sc[2], sc[3], sc[0], sc[1], engine=engine)
        for p in range(len(t_idx)):
            g, e = got[p], ef[p]
            gt = (int(g["score1"]), int(g["read_begin1"]), int(g["read_end1"]), int(g["ref_begin1"]), int(g["ref_end1"]))
            et = (int(e["score1"]), int(e["read_begin1"]), int(e["read_end1"]), int(e["ref_begin1"]), int(e["ref_end1"]))
            ok = gt == et if et[0] > 0 else gt[0] == 0
            # strand rule of ssw_check: forward iff score_f > score_r, ties -> '-'; mirrored coordinates
            m = int(lens[t_idx[p]])
            if int(ef[p]["score1"]) > int(er[p]["score1"]):
                x = ef[p]
                exp = (int(x["score1"]), int(x["read_begin1"]) + 1, int(x["read_end1"]) + 1, int(x["ref_begin1"]) + 1, int(x["ref_end1"]) + 1, 1)
            else:
                x = er[p]
                exp = (int(x["score1"]), m - int(x["read_end1"]), m - int(x["read_begin1"]), int(x["ref_begin1"]) + 1, int(x["ref_end1"]) + 1, -1)
            c = chk[p]
            gc = (int(c["score"]), int(c["qstart"]), int(c["qend"]), int(c["tstart"]), int(c["tend"]), int(c["strand"]))
            ok = ok and (gc == exp if exp[0] > 0 else gc[0] == 0)
            if not ok:
                print("MISMATCH", dict(scoring=sc, m=m, n=int(lens[r_idx[p]]), ssw=gt, oracle=et, check=gc, oracle_check=exp))
                np.savez("/tmp/soak_fail.npz", t=codes[int(offsets[t_idx[p]]):int(offsets[t_idx[p]]) + m],
                         r=codes[int(offsets[r_idx[p]]):int(offsets[r_idx[p]]) + int(lens[r_idx[p]])], sc=np.asarray(sc))
                sys.exit(1)
        n_pairs += len(t_idx); n_batches += 1; cells_total += cells
        print(f"batch {n_batches}: {len(t_idx)} pairs, scoring {sc}, {cells / 1e9:.0f} Gcells ok", flush=True)
    print(f"SOAK OK: {n_pairs} pairs in {n_batches} batches, {cells_total / 1e12:.2f} T forward cells per strand pair")


if __name__ == "__main__":
    main()
