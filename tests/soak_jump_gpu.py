#!/usr/bin/env python
"""Randomised parity soak of the sw_jump row (not collected by pytest: run by hand on a GPU box).

    python tests/soak_jump_gpu.py [--seconds 120] [--seed 1]

Draws batches of (contig, region1, region2) problems -- junction contigs with and without inserted bases between the two
parts, reversed part order (no valid jump), unrelated and low-complexity sequences, tandem repeats (many tied maxima), soft-masked
and N-rich stretches, region lengths on and around every boundary of the warp kernel's strip layout and beyond it (block
kernel), contigs shorter than a warp up to 4 kb -- with random scoring parameters, sends them through
smith_waterman.sw_jump_batch (nmsv_sw_jump_batch + nmsv_sw_jump_strings) and compares the complete return values, both
alignment strings included, with the C restatement of the reference's function (oracle/nmsv_jump_oracle.c, pinned on vectors
produced by nanomonsv/smith_waterman.py itself) computed on all host cores. Exits non-zero on the first difference.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nanomonsv_b200 import Engine, smith_waterman as SW   # noqa: E402
from oracle import oracle as O                            # noqa: E402

EDGES = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 767, 768, 769, 1023, 1024, 1025]


def rand_seq(rnd, n, kind):
    if kind == "low":
        return "".join(rnd.choice("AT") for _ in range(n))
    if kind == "tandem":
        unit = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 6)))
        return (unit * (n // len(unit) + 1))[:n]
    return "".join(rnd.choice("ACGT") for _ in range(n))


def mutate(rnd, s, rate):
    out = []
    for ch in s:
        r = rnd.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(rnd.choice("ACGT")); continue
        out.append(ch)
        if r > 1 - rate / 3:
            out.append(rnd.choice("ACGT"))
    return "".join(out) or "A"


def decorate(rnd, s):
    r = rnd.random()
    if r < 0.08 and len(s) > 8:          # soft-masked stretch: case matters to the reference (characters are compared)
        a = rnd.randrange(len(s) - 4); b = a + rnd.randint(1, len(s) - a)
        return s[:a] + s[a:b].lower() + s[b:]
    if r < 0.16 and len(s) > 8:          # N stretch: N == N is a match there
        a = rnd.randrange(len(s) - 4); b = min(len(s), a + rnd.randint(1, 30))
        return s[:a] + "N" * (b - a) + s[b:]
    return s


def region_len(rnd):
    r = rnd.random()
    if r < 0.25:
        return max(1, rnd.choice(EDGES) + rnd.choice([-1, 0, 0, 1]))
    if r < 0.30:
        return rnd.randint(1025, 2500)    # block kernel
    return rnd.randint(1, 1000)


def problem(rnd):
    kind = rnd.choice(["rand"] * 6 + ["low", "tandem"])
    m1, m2 = region_len(rnd), region_len(rnd)
    r1, r2 = rand_seq(rnd, m1, kind), rand_seq(rnd, m2, kind)
    shape = rnd.random()
    rate = rnd.choice([0.0, 0.01, 0.03, 0.08])
    flank = lambda: rand_seq(rnd, rnd.choice([0, 0, 3, rnd.randint(0, 300)]), kind)   # noqa: E731
    cut1, cut2 = rnd.randint(0, min(20, m1 - 1)), rnd.randint(0, min(20, m2 - 1))
    a, b = r1[cut1:m1 - rnd.randint(0, min(20, m1 - cut1 - 1))], r2[cut2:]
    if shape < 0.70:                      # junction: part of region1, optional inserted bases, part of region2
        ins = rand_seq(rnd, rnd.choice([0, 0, 0, 1, 5, 60, 200]), "rand")
        c = flank() + mutate(rnd, a, rate) + ins + mutate(rnd, b, rate) + flank()
    elif shape < 0.80:                    # wrong order: region2 part first
        c = flank() + mutate(rnd, b, rate) + mutate(rnd, a, rate) + flank()
    elif shape < 0.90:                    # unrelated
        c = rand_seq(rnd, rnd.randint(1, 1500), kind)
    else:                                 # contig shorter than a warp / tiny
        c = (a[:rnd.randint(1, 20)] + b[:rnd.randint(1, 20)]) or "A"
    if len(c) > 4000:
        c = c[:4000]
    return decorate(rnd, c), decorate(rnd, r1), decorate(rnd, r2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--batch", type=int, default=384)
    args = ap.parse_args()
    rnd = random.Random(args.seed)
    eng = Engine(0)
    workers = os.cpu_count() or 4
    pool = ThreadPoolExecutor(workers)      # the oracle call is a ctypes call: it releases the GIL
    t_end = time.time() + args.seconds
    n_prob = n_found = n_batches = 0
    cells = 0
    while time.time() < t_end:
        prm = rnd.choice([(1, 3, 3, 2, 10), (1, 2, 2, 2, 10), (1, 3, 3, 2, 10),
                          (rnd.randint(1, 3), rnd.randint(0, 4), rnd.randint(0, 4), rnd.randint(0, 5), rnd.randint(1, 30))])
        probs = [problem(rnd) for _ in range(args.batch)]
        got = SW.sw_jump_batch(probs, *prm, engine=eng)
        exp = list(pool.map(lambda p: O.sw_jump(p[0], p[1], p[2], *prm), probs))
        for k, (g, e) in enumerate(zip(got, exp)):
            if g != e:
                print("MISMATCH", prm, json.dumps(probs[k]), "\n got", g, "\n exp", e)
                sys.exit(1)
        n_prob += len(probs); n_found += sum(e is not None for e in exp); n_batches += 1
        cells += sum((len(c) + 1) * (len(a) + len(b) + 2) for c, a, b in probs)
    print(json.dumps({"soak": "sw_jump", "seed": args.seed, "seconds": args.seconds, "batches": n_batches, "problems": n_prob,
                      "found": n_found, "none": n_prob - n_found, "cells": cells, "mismatches": 0}))


if __name__ == "__main__":
    main()
