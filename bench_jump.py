#!/usr/bin/env python3
"""Secondary measurement (not the driver's headline bench): the sw_jump row (reference smith_waterman.py:5-127 /
locate_bp.py) on the GPU vs the C oracle on one host core, on locate_bp-shaped synthetic problems. Prints one JSON line.
A 'cell' is one H1 or H2 matrix entry; the reference's own pure-Python loop does ~8e4 cells/s (measured in the build
container on its __main__ example: 0.76 s for 700 x (46 + 43) cells)."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def problems(n, seed=7):
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        m1, m2 = rnd.randint(60, 900), rnd.randint(60, 900)
        r1 = "".join(rnd.choice("ACGT") for _ in range(m1))
        r2 = "".join(rnd.choice("ACGT") for _ in range(m2))
        c = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 400))) + r1[5:-5] + r2[5:-5] + \
            "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 400)))
        out.append((c, r1, r2))
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    import torch
    from nanomonsv_b200 import Engine, smith_waterman as SW
    from oracle import oracle as O
    eng = Engine(0)
    probs = problems(n)
    cells = sum((len(c) + 1) * (len(a) + len(b) + 2) for c, a, b in probs)
    SW.sw_jump_batch(probs[:64], 1, 3, 3, 2, 10, engine=eng)
    if os.environ.get("NMSV_JUMP_WARM"):      # grow the stream-ordered pool once (a long-running process has done so)
        SW.sw_jump_batch(probs, 1, 3, 3, 2, 10, engine=eng)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = SW.sw_jump_batch(probs, 1, 3, 3, 2, 10, engine=eng)
    dt = time.perf_counter() - t0
    dt_abi = SW.sw_jump_batch.last_abi_seconds
    k = min(40, n)
    t1 = time.perf_counter()
    exp = [O.sw_jump(c, a, b, 1, 3, 3, 2, 10) for c, a, b in probs[:k]]
    dt_cpu = time.perf_counter() - t1
    assert got[:k] == exp
    cells_k = sum((len(c) + 1) * (len(a) + len(b) + 2) for c, a, b in probs[:k])
    print(json.dumps({"metric": "sw_jump cells/s (end to end through nmsv_sw_jump_batch, host strings in and out)",
                      "problems": n, "cells": cells, "gpu_seconds": dt, "gpu_gcells_per_s": cells / dt / 1e9,
                      "abi_seconds": dt_abi, "abi_gcells_per_s": cells / dt_abi / 1e9,
                      "problems_per_s": n / dt, "found": sum(g is not None for g in got),
                      "cpu_port_1core_gcells_per_s": cells_k / dt_cpu / 1e9,
                      "reference_python_gcells_per_s": 8e-5, "device": eng.device_info()["name"]}))


if __name__ == "__main__":
    main()
