"""Sizes of the begin-pass tasks of the bench batch (on a GPU box): window lengths, tiles, warp-steps per slot."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from types import SimpleNamespace
import bench
from nanomonsv_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
cfg, pack, svs, reads = bench.build_batch(n, 0, SimpleNamespace())
eng = Engine(0)
plan = eng.plan(pack, svs, reads); plan.set_profiling(True)
for _ in range(2): plan.run()
torch.cuda.synchronize()
plan.download()
print(plan.kernel_times()); print(plan.stats())
alns, flags = plan.download_alns()
cand = (flags & 1) != 0
print("reads", len(flags), "candidates", int(cand.sum()))
lens = np.asarray(pack._lens)
steps = []
for r in np.nonzero(cand)[0]:
    sv = svs[reads[r]["sv"]]
    seen = set()
    for s in range(4):
        t = int(sv["tmpl"][s])
        if t == 0xffffffff or t in seen: continue
        seen.add(t)
        a = alns[r, s]
        if a["score"] <= 0: continue
        m = int(lens[t])
        rows = int(a["qend"]) if a["strand"] > 0 else m - int(a["qstart"]) + 1
        ecol = int(a["tend"]) - 1
        w = min(ecol + 1, rows + (2 * rows - int(a["score"])))
        tiles = (rows + 511) // 512
        span = int(a["tend"]) - int(a["tstart"]) + 1
        w1 = rows + rows // 8 + 64
        steps.append((tiles * (w + 31), tiles, w, rows, int(a["score"]), s, span, span > w1))
steps.sort(reverse=True)
tot = sum(x[0] for x in steps)
print("begin tasks", len(steps), "total warp-steps", tot, "max", steps[:5], "mean", tot / max(len(steps), 1))
by_slot = {}
for x in steps: by_slot.setdefault(x[5], [0, 0]); by_slot[x[5]][0] += 1; by_slot[x[5]][1] += x[0]
print("by slot", by_slot)

over = [x for x in steps if x[7]]
print("spans beyond first window", len(over), over[:10])
