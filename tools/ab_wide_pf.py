#!/usr/bin/env python
"""CTA-per-world fused kernel with / without the batched L2 state prefetch (MAPF_DBG_FLAGS bit 29 = off), ms per step."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, generate_scenario_device
from primal_ppo_b200.build import build
build()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
for (W, H, N, F) in ((16384, 80, 128, 9), (8192, 80, 128, 15), (4096, 80, 128, 21), (2048, 80, 128, 31), (32768, 40, 48, 9), (65536, 40, 33, 9), (8192, 128, 128, 9)):
    dsc = generate_scenario_device(W, H, H, N, kind="density", density=(0.0, 0.3), queue_len=4, seed=3, device=dev, fov=F)
    obs = torch.empty((W, N, 6, F, F), device=dev); vec = torch.empty((W, N, 4), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(4)]
    row = {"worlds": W, "grid": H, "agents": N, "fov": F}
    for rep in range(2):
        for name, flags in (("prefetch", 0), ("off", 1 << 29)):
            os.environ["MAPF_DBG_FLAGS"] = str(flags)
            env = BatchedMapfGym(dsc, device=dev, use_tape=False)
            for i in range(4):
                env.step_observe(ring[i % 4], obs_out=(obs, vec))
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20):
                env.step_observe(ring[i % 4], obs_out=(obs, vec))
            b.record(); torch.cuda.synchronize()
            row[f"{name} #{rep}"] = round(a.elapsed_time(b) / 20, 4)
            del env
    print(json.dumps(row), flush=True)
    del obs, vec, ring, dsc
    torch.cuda.empty_cache()
