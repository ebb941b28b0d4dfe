"""Three calls of nmsv_sw_jump_batch on the 3000 problems of bench_jump.py (for the ncu capture of sw_jump_warp_kernel:
the third launch is the warm one)."""
import os, sys
sys.path.insert(0, os.getcwd())
import bench_jump
from nanomonsv_b200 import Engine, smith_waterman as SW
probs = bench_jump.problems(int(sys.argv[1]) if len(sys.argv) > 1 else 3000)
eng = Engine(0)
for _ in range(3):
    SW.sw_jump_batch(probs, 1, 3, 3, 2, 10, engine=eng)
print("abi ms", SW.sw_jump_batch.last_abi_seconds * 1e3, file=sys.stderr)
