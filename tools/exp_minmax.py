#!/usr/bin/env python
"""Per-launch timing distribution of mapf_observe / mapf_step_observe (CUDA events around every launch)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, random_scenario
W, N = 65536, 32
sc = random_scenario(W, 40, 40, N, density=(0.0, 0.3), queue_len=16, seed=100, unique_maps=256)
env = BatchedMapfGym(sc, use_tape=False)
dev = env.device
obs = torch.empty((W, N, 6, 9, 9), device=dev); vec = torch.empty((W, N, 4), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]
big = torch.empty(1 << 30, dtype=torch.float32, device=dev)
for name, fn in (("observe", lambda i: env.getAllObservations(out=(obs, vec))), ("fused", lambda i: env.step_observe(ring[i % 8], obs_out=(obs, vec))),
                 ("torch fill 4 GiB", lambda i: big.zero_())):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    K = 40
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        ev[i][0].record(); fn(i); ev[i][1].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev])
    print(f"{name:18s} min {t.min():.4f} median {np.median(t):.4f} mean {t.mean():.4f} max {t.max():.4f} ms", flush=True)
    # isolated launches (a sync and a short pause between them)
    iso = []
    for i in range(10):
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record(); torch.cuda.synchronize(); iso.append(a.elapsed_time(b))
    print(f"{name:18s} isolated: min {min(iso):.4f} median {np.median(iso):.4f} ms", flush=True)
