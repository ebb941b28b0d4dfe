"""2-rank NCCL test of the product's multi-GPU path (skipped on a box with fewer than 2 GPUs): the SV candidates of one
job are partitioned with shard.lpt_partition, each rank validates its share on ITS GPU through the C ABI, counts are
merged by one all-reduce and the records by the count all-gather + padded all-gather on the device; the merged result
must equal a single-GPU run of the whole job."""
import dataclasses
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _job():
    from nanomonsv_b200 import synth
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=12, coverage=8.0, read_len_mean=2500,
                              read_len_min=600, read_len_max=7000, contig_len=150000, n_contigs=2, seed=97,
                              short_fraction=0.6)
    wl = synth.make_workload(cfg, samples=("tumor",))
    return wl.to_batch("tumor", var_read_min_mapq=0)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from bench_strong import _sub_batch
    from nanomonsv_b200 import Engine, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pack, svs, reads = _job()
    arrays = pack.finalize()
    costs = shard.sv_costs(svs, reads, arrays[2])
    eng = Engine(rank)

    def run_local(picks):
        spack, ssvs, sreads = _sub_batch(arrays, svs, reads, picks)
        checked, supporting, records = eng.validate(spack, ssvs, sreads)
        return checked, supporting, shard.records_to_rows(records, ssvs, picks)

    table, rows, parts = shard.validate_sharded(costs, run_local, rank, world, device=dev)
    if rank == 0:
        c1, s1, r1 = eng.validate(pack, svs, reads)
        rows1 = shard.records_to_rows(r1, svs, list(range(len(svs))))
        rows1 = rows1[np.lexsort((rows1[:, 1], rows1[:, 0]))]
        out.put((bool((table[:, 0] == c1).all() and (table[:, 1] == s1).all()), bool((rows == rows1).all()), len(rows),
                 [len(p) for p in parts]))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_nccl_merge_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    counts_ok, rows_ok, n_rows, sizes = out.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counts_ok and rows_ok and n_rows >= 5 and all(sizes)
