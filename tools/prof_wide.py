#!/usr/bin/env python
"""Short driver for ncu: BASELINE configs[4] at FOV 9 (80x80 worlds, 128 agents) through the CTA-per-world fused kernel
(step_observe_wide_kernel), then all BFS maps of 8192 40x40x32 worlds (bfs_gray_kernel) and the two-stream GAE."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, gae, generate_scenario_device  # noqa: E402
from primal_ppo_b200.build import build  # noqa: E402

build()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
W, H, N, F = 16384, 80, 128, 9
dsc = generate_scenario_device(W, H, H, N, kind="density", density=(0.0, 0.3), queue_len=8, seed=909, device=dev, fov=F)
env = BatchedMapfGym(dsc, device=dev, use_tape=False)
obs = torch.empty((W, N, 6, F, F), device=dev); vec = torch.empty((W, N, 4), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
for i in range(6):
    a = torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8)
    env.step_observe(a, obs_out=(obs, vec))
torch.cuda.synchronize()
del env, obs, vec, dsc
torch.cuda.empty_cache()
W, H, N = 8192, 40, 32
dsc = generate_scenario_device(W, H, H, N, kind="density", density=(0.0, 0.3), queue_len=8, seed=910, device=dev)
env = BatchedMapfGym(dsc, device=dev, use_tape=False)
maps = env.bfs_maps()
for i in range(3):
    env.bfs_maps(out=maps)
T, cols = 256, 8192 * 32
r, v, lv = torch.randn((T, cols), device=dev), torch.randn((T, cols), device=dev), torch.randn((cols,), device=dev)
for i in range(3):
    gae(r, v, lv)
torch.cuda.synchronize()
print("ok")
