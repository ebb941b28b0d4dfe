#!/usr/bin/env python
"""Randomised parity soak of the STAGE seam (not collected by pytest: run by hand on a GPU box).

    python tests/soak_validate_gpu.py [--seconds 180] [--seed 1]

Seeded synthetic workloads with random shapes (canonical / single-breakend, validate_sequence_length 100 - 1500, share
of short deletions / duplications, coverage, read lengths, error rates, mapq thresholds, score ratios) go through
count_sread_from_seam on the GPU and through the CPU oracle's restatement of Alignment_counter; the sread_count and
sread_info files must be identical. Exits non-zero on the first difference.
"""
from __future__ import annotations

import argparse
import dataclasses
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nanomonsv_b200 import count_sread_by_alignment as C   # noqa: E402
from nanomonsv_b200 import Engine, synth                   # noqa: E402
from oracle import oracle as O                             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    eng = Engine(0)
    t_end = time.time() + args.seconds
    n_runs = n_sv = n_supp = 0
    tmp = tempfile.mkdtemp(prefix="nmsv_soak_")
    while time.time() < t_end:
        sbnd = bool(rng.random() < 0.3)
        L = int(rng.choice([100, 200, 200, 300, 700, 1000, 1500]))
        mean = float(rng.choice([1500, 3000, 8000]))
        cfg = dataclasses.replace(
            synth.PRESETS["single_breakend" if sbnd else "testdata_shaped"], n_sv=int(rng.integers(4, 14)),
            validate_sequence_length=L, coverage=float(rng.choice([6.0, 12.0, 20.0])), read_len_mean=mean,
            read_len_min=int(rng.choice([200, 600, 1000])), read_len_max=int(mean * rng.choice([2, 4])),
            contig_len=200000, n_contigs=2, seed=int(rng.integers(1, 2 ** 31)),
            **({} if sbnd else {"short_fraction": float(rng.choice([0.0, 0.3, 0.8]))}))
        wl = synth.make_workload(cfg)
        min_mapq = int(rng.choice([0, 20, 40]))
        ratio = float(rng.choice([1.0, 1.2, 1.4]))
        for sample in ("tumor", "control"):
            seam = os.path.join(tmp, f"{n_runs}.{sample}.tsv")
            wl.write_seam(sample, seam)
            O.count_from_seam(seam, wl.fasta, seam + ".o_count", seam + ".o_info", is_sbnd=sbnd, var_read_min_mapq=min_mapq,
                              score_ratio_thres=ratio, validate_sequence_length=L, engine="striped")
            C.count_sread_from_seam(seam, seam + ".g_count", seam + ".g_info", wl.fasta, sbnd, min_mapq, ratio, engine=eng,
                                    validate_sequence_length=L)
            ec, gc = open(seam + ".o_count").read(), open(seam + ".g_count").read()
            ei, gi = sorted(open(seam + ".o_info").read().splitlines()), sorted(open(seam + ".g_info").read().splitlines())
            if ec != gc or ei != gi:
                print("MISMATCH", dict(sbnd=sbnd, L=L, seed=cfg.seed, sample=sample, min_mapq=min_mapq, ratio=ratio), "files kept in", tmp)
                sys.exit(1)
            n_supp += len(gi)
            for f in (seam, seam + ".o_count", seam + ".o_info", seam + ".g_count", seam + ".g_info"):
                os.remove(f)
        n_runs += 1; n_sv += cfg.n_sv
        print(f"run {n_runs}: {'sbnd' if sbnd else 'canonical'} L={L} n_sv={cfg.n_sv} ok", flush=True)
    print(f"SOAK OK: {n_runs} workloads, {n_sv} SV candidates x 2 samples, {n_supp} supporting-read lines identical")


if __name__ == "__main__":
    main()
