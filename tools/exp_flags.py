#!/usr/bin/env python
"""Times mapf_step_observe / mapf_step / mapf_observe at 65 536 x 40x40 x 32 under the current MAPF_DBG_FLAGS
(kernel experiment switches).  Usage: for f in 240 16 0 64; do MAPF_DBG_FLAGS=$f python tools/exp_flags.py; done"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, random_scenario  # noqa: E402

W, N = int(os.environ.get("W", 65536)), 32
sc = random_scenario(W, 40, 40, N, density=(0.0, 0.3), queue_len=16, seed=100, unique_maps=256)
env = BatchedMapfGym(sc, use_tape=False)
dev = env.device
obs = torch.empty((W, N, 6, 9, 9), device=dev)
vec = torch.empty((W, N, 4), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]


def timeit(fn, K=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K


t_f = timeit(lambda i: env.step_observe(ring[i % 8], obs_out=(obs, vec)))
t_s = timeit(lambda i: env.step(ring[i % 8]))
t_o = timeit(lambda i: env.getAllObservations(out=(obs, vec)))
t_2 = timeit(lambda i: (env.step(ring[i % 8]), env.getAllObservations(out=(obs, vec))))
print(f"flags={os.environ.get('MAPF_DBG_FLAGS', '0'):>5s} fused={t_f:.4f} step={t_s:.4f} observe={t_o:.4f} two={t_2:.4f} ms", flush=True)
