#!/usr/bin/env python
"""A small end-to-end exercise of every kernel, sized for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, gae, generate_scenario_device, random_actions, random_scenario, sample_actions

for (W, H, Wd, N, F, da) in ((40, 8, 8, 8, 9, False), (9, 33, 65, 5, 9, False), (6, 20, 20, 48, 15, True), (5, 40, 40, 32, 9, False)):
    sc = random_scenario(W, H, Wd, N, density=(0.15, 0.3), queue_len=3, seed=W, fov=F, use_da=da, use_hp=da)
    env = BatchedMapfGym(sc, use_tape=False)
    acts = torch.from_numpy(random_actions(6, W, N, seed=1)).cuda()
    maps = env.bfs_maps()
    for t in range(6):
        if t % 2 == 0:
            out, obs, vec = env.step_observe(acts[t])
        else:
            st = env.getActionStatus(acts[t]); env.getTrainValid(acts[t]); env.jointStep(acts[t], st); obs, vec = env.getAllObservations()
            out = env.step(acts[t])
        env.refresh_bfs(maps)
    ob16 = torch.empty(obs.shape, dtype=torch.bfloat16, device="cuda")
    env.getAllObservations(out=(ob16, vec))
    env.counters(); env.state()
    hb = env.make_host_buffers()
    env.step_observe_host(hb, obs, vec)
d = generate_scenario_device(64, 40, 60, 6, kind="warehouse", queue_len=3, seed=2, human_loops=2)
d2 = generate_scenario_device(64, 24, 24, 6, kind="density", density=(0.0, 0.3), triangular=True, size_range=(10, 24), queue_len=3, seed=3)
BatchedMapfGym(d, use_tape=False).step_observe(torch.zeros((64, 6), dtype=torch.int8))
ps = torch.softmax(torch.randn(1000, 5, device="cuda"), -1)
sample_actions(ps, seed=1, draw=2)
gae(torch.randn(16, 1000, device="cuda"), torch.randn(16, 1000, device="cuda"), torch.randn(1000, device="cuda"))
gae(torch.randn(7, 5, device="cuda"), torch.randn(7, 5, device="cuda"), torch.randn(5, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
