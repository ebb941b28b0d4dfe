"""BASELINE config 5 alone (bench.pair_sweep_entry: 10^6 pairs through nmsv_ssw_check_batch, serial and two calls in flight)."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from nanomonsv_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device("cuda", 0)
eng = Engine(0)
e = bench.pair_sweep_entry(torch, None, dev, eng, 0, 1, n, 80.0, 1.0, os.cpu_count())
print(json.dumps(e))
