#!/usr/bin/env python
"""In-process A/B of kernel experiment switches: one env per MAPF_DBG_FLAGS value (the flags are read at mapf_create),
all writing the SAME observation tensors, measured in interleaved rounds — process-to-process placement noise (+-3 %)
cancels.    python tools/ab_inproc.py 0 240 16 128"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, random_scenario
flags = [int(x) for x in sys.argv[1:]] or [0, 240]
W, N = 65536, 32
sc = random_scenario(W, 40, 40, N, density=(0.0, 0.3), queue_len=16, seed=100, unique_maps=256)
envs = {}
for f in flags:
    os.environ["MAPF_DBG_FLAGS"] = str(f)
    envs[f] = BatchedMapfGym(sc, use_tape=False)
dev = envs[flags[0]].device
obs = torch.empty((W, N, 6, 9, 9), device=dev); vec = torch.empty((W, N, 4), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]
res = {f: {"fused": [], "observe": [], "step": []} for f in flags}
for rnd in range(6):
    for f in flags:
        e = envs[f]
        for name, fn in (("fused", lambda i: e.step_observe(ring[i % 8], obs_out=(obs, vec))),
                         ("observe", lambda i: e.getAllObservations(out=(obs, vec))), ("step", lambda i: e.step(ring[i % 8]))):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20):
                fn(i)
            b.record(); torch.cuda.synchronize()
            if rnd:
                res[f][name].append(a.elapsed_time(b) / 20)
for f in flags:
    print(f"flags={f:6d} " + " ".join(f"{k}={np.median(v):.4f}" for k, v in res[f].items()), flush=True)
