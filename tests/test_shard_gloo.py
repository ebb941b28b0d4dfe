"""CPU tests of the multi-GPU host logic: partitioning and the final count merge, run as 2 real processes over the
gloo backend (NCCL takes its place on the GPU box). The per-rank engine is replaced by the CPU oracle here -- tests may
do that, the product path may not."""
import dataclasses
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanomonsv_b200 import shard, synth
from oracle import oracle as O


def test_lpt_partition_is_balanced_and_complete():
    rng = np.random.default_rng(1)
    costs = rng.gamma(2.0, 5.0, size=203)
    for n in (1, 2, 4, 8):
        parts = shard.lpt_partition(costs, n)
        assert sorted(i for p in parts for i in p) == list(range(203))
        loads = [costs[p].sum() for p in parts]
        assert max(loads) - min(loads) <= costs.max() + 1e-9
    assert shard.round_robin_partition(7, 3) == [[0, 3, 6], [1, 4], [2, 5]]
    assert shard.lpt_partition([], 4) == [[], [], [], []]


def test_sv_costs_counts_shared_templates_once():
    cfg = synth.SynthConfig(n_sv=6, contig_len=120000, n_contigs=2, read_len_mean=1500, read_len_min=600,
                            read_len_max=3000, coverage=4.0, seed=3, validate_sequence_length=200)
    wl = synth.make_workload(cfg)
    pack, svs, reads = wl.to_batch("tumor")
    codes, offsets, lens = pack.finalize()
    costs = shard.sv_costs(svs, reads, lens)
    for i, sv in enumerate(svs):
        bases = int(lens[reads["seq"][sv["read_begin"]:sv["read_end"]]].sum())
        distinct = {int(t) for t in sv["tmpl"] if t != 0xFFFFFFFF}
        assert costs[i] == 2 * bases * sum(int(lens[t]) for t in distinct)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rows_of(i, seam_rows_i, info_lines):
    """Oracle info lines of SV i -> the int32 record rows that the GPU path gathers (shard.RECORD_COLS)."""
    ids = [r[0] for r in seam_rows_i]
    rows = []
    for line in info_lines:
        f = line.split("\t")
        vals = [0 if x == "None" else (1 if x == "+" else (-1 if x == "-" else int(x))) for x in f[9:]]
        rows.append([i, ids.index(f[8])] + vals)
    return np.asarray(rows, dtype=np.int32).reshape(-1, shard.RECORD_COLS)


def _worker(rank, world, port, seam_rows, contigs, keys, costs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fasta = O.DictFasta(contigs)

    def run_local(indices):
        checked, supporting, rows = [], [], [np.zeros((0, shard.RECORD_COLS), dtype=np.int32)]
        for i in indices:
            count_line, info = O.count_one_sv(keys[i], seam_rows[i], fasta, var_read_min_mapq=0, engine="striped")
            f = count_line.split("\t")
            checked.append(int(f[-2])); supporting.append(int(f[-1])); rows.append(_rows_of(i, seam_rows[i], info))
        return checked, supporting, np.concatenate(rows)

    table, rows, parts = shard.validate_sharded(costs, run_local, rank, world)
    if rank == 0:
        out.put((table.tolist(), rows.tolist(), parts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_merge_equals_single_process():
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=7, coverage=4.0, read_len_mean=1200,
                              read_len_min=500, read_len_max=2500, contig_len=60000, n_contigs=2, seed=21)
    wl = synth.make_workload(cfg, samples=("tumor",))
    keys = [s.key for s in wl.svs]
    seam_rows = [[(r.rid, str(r.mapq1), str(r.mapq2), synth.codes_to_str(r.codes)) for r in per]
                 for per in wl.samples["tumor"]]
    contigs = {k: synth.codes_to_str(v) for k, v in wl.fasta.contigs.items()}
    costs = [sum(len(r[3]) for r in rows) for rows in seam_rows]
    # single process truth
    fasta = O.DictFasta(contigs)
    exp_table, exp_rows = [], []
    for i, (k, rows) in enumerate(zip(keys, seam_rows)):
        cl, info = O.count_one_sv(k, rows, fasta, var_read_min_mapq=0, engine="striped")
        exp_table.append([int(cl.split("\t")[-2]), int(cl.split("\t")[-1])])
        exp_rows += _rows_of(i, rows, info).tolist()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seam_rows, contigs, keys, costs, out)) for r in range(2)]
    for p in procs:
        p.start()
    table, rows, parts = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert table == exp_table
    # records of all ranks, gathered with a count all-gather + one padded all-gather, sorted by (SV, read)
    assert len(exp_rows) >= 3 and rows == sorted(exp_rows, key=lambda r: (r[0], r[1]))
    assert sorted(i for p in parts for i in p) == list(range(7)) and all(parts)
