#!/usr/bin/env python
"""Fused step+observe over a range of shapes: ms per step, algorithmic GB/s and fraction of the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from primal_ppo_b200 import BatchedMapfGym, generate_scenario_device  # noqa: E402
from primal_ppo_b200.build import build  # noqa: E402

build()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
for (W, H, N, F) in ((4096, 20, 8, 9), (65536, 20, 8, 9), (262144, 10, 8, 9), (131072, 20, 16, 9), (262144, 20, 2, 9), (131072, 40, 4, 9), (65536, 40, 12, 9),
                     (65536, 40, 32, 9)):
    dsc = generate_scenario_device(W, H, H, N, kind="density", density=(0.0, 0.3), queue_len=4, seed=3, device=dev, fov=F)
    env = BatchedMapfGym(dsc, device=dev, use_tape=False)
    obs = torch.empty((W, N, 6, F, F), device=dev); vec = torch.empty((W, N, 4), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(4)]
    for i in range(5):
        env.step_observe(ring[i % 4], obs_out=(obs, vec))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    a.record()
    for i in range(K):
        env.step_observe(ring[i % 4], obs_out=(obs, vec))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    _, _, bf = bench.algorithmic_bytes(N, H, H, 6, F)
    gbs = bf * W * N / (ms * 1e-3) / 1e9
    print(json.dumps({"worlds": W, "grid": H, "agents": N, "fov": F, "ms": round(ms, 4), "agent_steps_per_s": round(W * N / (ms * 1e-3) / 1e9, 3),
                      "GB/s": round(gbs), "frac": round(gbs / peak, 3), "obs_GB": round(W * N * 24 * F * F / 1e9, 2)}), flush=True)
    del env, obs, vec, ring, dsc
    torch.cuda.empty_cache()
