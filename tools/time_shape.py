"""One device-resident run of a secondary workload shape (for ncu captures): python tools/time_shape.py <preset> <n_sv>"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from nanomonsv_b200 import Engine
preset, n = sys.argv[1], int(sys.argv[2])
cfg, pack, svs, reads = bench.build_batch(n, 0, None, preset=preset, budget=None, contig_len=2_000_000, n_contigs=2)
eng = Engine(0)
plan = eng.plan(pack, svs, reads); plan.set_profiling(True)
for _ in range(2): plan.run()
torch.cuda.synchronize()
kt = plan.kernel_times(); plan.download(); st = plan.stats()
ms = sum(t for n_, r, t in kt if n_ == "sw_forward")
print(preset, n, "fwd_ms", round(ms, 2), "gcups_exec", round(st["cells_executed"] / ms / 1e6), file=sys.stderr)
