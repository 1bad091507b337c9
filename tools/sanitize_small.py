#!/usr/bin/env python
"""A small end-to-end exercise of every kernel, sized for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import (BatchedMapfGym, StepOut, checksum_rows, decode_results, gae, gae2, generate_scenario_device,
                             random_actions, random_scenario, sample_actions)

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
# round 2: goal sampling on device, the CTA-per-world fused kernel (4- and 8-warp plans), split-phase host calls (full and
# compact slab), packed results, allGoodActions, row checksums, two-stream GAE, the fused PPO-loss kernel
for (W, H, N, F) in ((12, 24, 48, 9), (6, 40, 128, 9), (4, 40, 128, 31), (16, 12, 8, 9), (5, 40, 32, 31)):
    sc = random_scenario(W, H, H, N, density=(0.1, 0.25), queue_len=1, seed=W + N, fov=F)
    env = BatchedMapfGym(sc, use_tape=False, goal_sampling=True)
    acts = random_actions(6, W, N, seed=2)
    obs = torch.empty((W, N, 6, F, F), device="cuda"); vec = torch.empty((W, N, 4), device="cuda")
    for compact in (False, True):
        ring = env.make_host_ring(slots=2, action_slots=2, compact=compact, with_train_valid=not compact)
        tv = torch.empty((W, N, 5), device="cuda")
        for t in range(4):
            ring["action_ring"][t & 1].copy_(torch.from_numpy(acts[t]))
            env.step_observe_host_begin(ring["action_ring"][t & 1], ring["slots"][t & 1], obs, vec,
                                        train_valid_dev=None if compact else tv, with_train_valid=not compact)
            if t:
                env.host_wait(1)
        env.host_wait(0)
        if compact:
            decode_results(ring["slots"][1]["packed"])
    packed = torch.empty((W, N), dtype=torch.int16, device="cuda")
    o = env._out
    env.step(torch.from_numpy(acts[4]), out=StepOut(status=o.status, reward=o.reward, cost=o.cost, train_valid=o.train_valid,
                                                     goals_reached=o.goals_reached, violated=o.violated, shadow_goals=o.shadow_goals,
                                                     fixed_actions=o.fixed_actions, packed=packed))
    env.step_observe(torch.from_numpy(acts[5]), obs_out=(obs, vec))
    _ = env.allGoodActions, env.human(), checksum_rows(obs)
r = [torch.randn(9, 1000, device="cuda") for _ in range(4)]
gae2(r[0], r[1], torch.randn(1000, device="cuda"), r[2], r[3], torch.randn(1000, device="cuda"))
from primal_ppo_b200.ppo.fused_loss import fused_ppo_lagrange_loss
from primal_ppo_b200.ppo.loss import PPOConfig
from primal_ppo_b200.ppo.policy import PolicyOutput
B, N = 7, 5
kw = {f: None for f in PolicyOutput._fields}
kw.update(policy=torch.softmax(torch.randn(B, N, 5, device="cuda"), -1).requires_grad_(True), value=torch.randn(B, N, 1, device="cuda", requires_grad=True),
          cost_value=torch.randn(B, N, 1, device="cuda", requires_grad=True), policy_sig=torch.sigmoid(torch.randn(B, N, 5, device="cuda")).requires_grad_(True))
loss, _ = fused_ppo_lagrange_loss(PolicyOutput(**kw), returns=torch.randn(B, N, device="cuda"), cost_returns=torch.randn(B, N, device="cuda"),
                                  old_v=torch.randn(B, N, device="cuda"), old_cv=torch.randn(B, N, device="cuda"),
                                  actions=torch.randint(0, 5, (B, N), device="cuda", dtype=torch.int8),
                                  old_ps=torch.softmax(torch.randn(B, N, 5, device="cuda"), -1),
                                  train_valid=torch.ones(B, N, 5, device="cuda"), lagrangian=0.5, cfg=PPOConfig(cost_coef=0.1))
loss.backward()
d = generate_scenario_device(64, 40, 60, 6, kind="warehouse", queue_len=3, seed=2, human_loops=2)
d2 = generate_scenario_device(64, 24, 24, 6, kind="density", density=(0.0, 0.3), triangular=True, size_range=(10, 24), queue_len=3, seed=3)
BatchedMapfGym(d, use_tape=False).step_observe(torch.zeros((64, 6), dtype=torch.int8))
ps = torch.softmax(torch.randn(1000, 5, device="cuda"), -1)
sample_actions(ps, seed=1, draw=2)
gae(torch.randn(16, 1000, device="cuda"), torch.randn(16, 1000, device="cuda"), torch.randn(1000, device="cuda"))
gae(torch.randn(7, 5, device="cuda"), torch.randn(7, 5, device="cuda"), torch.randn(5, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
