#!/usr/bin/env python
"""Timing probe of the sw_jump warp kernel (NMSV_TRACE prints the kernel time): uniform-width problem sets vs the mixed set
of bench_jump.py.   python tools/jump_probe.py mixed|uniform [n]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_jump
from nanomonsv_b200 import Engine, smith_waterman as SW
kind = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
rnd = random.Random(5)
if kind == "mixed":
    probs = bench_jump.problems(n)
else:
    probs = []
    for _ in range(n):
        m1 = m2 = 500
        if kind == "pow2": m1, m2 = rnd.choice([64, 128, 256, 512, 1024]), rnd.choice([64, 128, 256, 512, 1024])
        if kind == "three": m1, m2 = rnd.choice([256, 512, 1024]), rnd.choice([256, 512, 1024])
        if kind == "even": m1, m2 = 64 * rnd.randint(1, 16), 64 * rnd.randint(1, 16)
        r1 = "".join(rnd.choice("ACGT") for _ in range(m1)); r2 = "".join(rnd.choice("ACGT") for _ in range(m2))
        c = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 400))) + r1[5:-5] + r2[5:-5] + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 400)))
        probs.append((c, r1, r2))
cells = sum((len(c) + 1) * (len(a) + len(b) + 2) for c, a, b in probs)
eng = Engine(0)
MS = int(os.environ.get('PROBE_MS', '10'))
SW.sw_jump_batch(probs, 1, 3, 3, 2, MS, engine=eng)
SW.sw_jump_batch(probs, 1, 3, 3, 2, MS, engine=eng)
print(kind, n, "cells %.3e" % cells, "abi s", SW.sw_jump_batch.last_abi_seconds, "Gcells/s %.1f" % (cells / SW.sw_jump_batch.last_abi_seconds / 1e9))
