"""Forward-launch time of the bench batch for a given build (on a GPU box):

    NMSV_LIB_PATH=build/variants/lib_x.so NMSV_FUSED_BEGIN=0 python tools/time_kernel.py 192

Prints the best of three runs of the sw_forward launch (CUDA events inside the library) and the executed GCUPS; with
NMSV_FUSED_BEGIN=0 the launch holds the forward work only. Variant libraries: tools/mk_variant.sh name [-Dflags].
"""
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
from types import SimpleNamespace
import bench
from nanomonsv_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cfg, pack, svs, reads = bench.build_batch(n, 0, SimpleNamespace())
eng = Engine(0)
plan = eng.plan(pack, svs, reads); plan.set_profiling(True)
for _ in range(3): plan.run()
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    plan.run(); torch.cuda.synchronize()
    kt = plan.kernel_times()
    best = min(best, [ms for n_, r, ms in kt if n_ == "sw_forward"][0])
plan.download()
st = plan.stats()
print({k: st[k] for k in ("cells_reference", "cells_dedup", "cells_executed", "cells_pruned", "cells_begin", "fwd_tasks", "fwd_tasks_queued", "rev_tasks", "reads_pruned")}, file=sys.stderr)
print(os.environ.get("NMSV_PRUNE"), os.environ.get("NMSV_SPAWN_FIRST"), n, os.environ.get("NMSV_LIB_PATH"), "fwd_ms", round(best, 2), "gcups_exec", round(st["cells_executed"] / best / 1e6))
