#!/usr/bin/env python
"""Times mapf_bfs (all maps of 8192 worlds 40x40x32) under the current MAPF_DBG_FLAGS."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, random_scenario
W, N, H = 8192, 32, int(os.environ.get("H", 40))
sc = random_scenario(W, H, H, N, density=(0.0, 0.3), queue_len=2, seed=100, unique_maps=128)
env = BatchedMapfGym(sc, use_tape=False)
out = env.bfs_maps(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    env.bfs_maps(out=out)
b.record(); torch.cuda.synchronize()
print(f"flags={os.environ.get('MAPF_DBG_FLAGS','0'):>4s} H={H} bfs {a.elapsed_time(b)/5:.3f} ms for {W*N} maps  checksum {int(out.long().sum())}", flush=True)
