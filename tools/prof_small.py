#!/usr/bin/env python
"""Short driver for ncu: the fused step+observe launch on small worlds (lane groups): 65 536 x 20x20x8 and 131 072 x 20x20x16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, generate_scenario_device  # noqa: E402
from primal_ppo_b200.build import build  # noqa: E402

build()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
for (W, H, N) in ((65536, 20, 8), (131072, 20, 16)):
    dsc = generate_scenario_device(W, H, H, N, kind="density", density=(0.0, 0.3), queue_len=4, seed=3, device=dev)
    env = BatchedMapfGym(dsc, device=dev, use_tape=False)
    obs = torch.empty((W, N, 6, 9, 9), device=dev); vec = torch.empty((W, N, 4), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    for i in range(4):
        a = torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8)
        env.step_observe(a, obs_out=(obs, vec))
    torch.cuda.synchronize()
    del env, obs, vec, dsc
    torch.cuda.empty_cache()
print("ok")
