#!/usr/bin/env python
"""Experiment: where does the fused step+observe launch spend its time?  Times mapf_step_observe / mapf_observe /
mapf_step with subsets of the step outputs disabled (NULL pointers), CUDA events, 65 536 x 40x40 x 32."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import BatchedMapfGym, _cabi, random_scenario  # noqa: E402

W, N = int(os.environ.get("W", 65536)), 32
sc = random_scenario(W, 40, 40, N, density=(0.0, 0.3), queue_len=16, seed=100, unique_maps=256)
env = BatchedMapfGym(sc, use_tape=False)
dev = env.device
obs = torch.empty((W, N, 6, 9, 9), device=dev)
vec = torch.empty((W, N, 4), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]
lib = env._lib
stream = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def so(keys):
    o = env._out
    d = {k: (getattr(o, k).data_ptr() if k in keys else None) for k in
         ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow_goals", "fixed_actions")}
    return _cabi.MapfStepOut(**d)


ALL = ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow_goals", "fixed_actions")


def timeit(name, fn, K=30):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        fn(i)
    b.record(); torch.cuda.synchronize()
    print(f"{name:58s} {a.elapsed_time(b) / K:.4f} ms", flush=True)


for label, keys in (("all outputs", ALL), ("no train_valid", tuple(k for k in ALL if k != "train_valid")),
                    ("reward only", ("reward",)), ("no outputs", ())):
    s = so(keys)
    timeit(f"fused step_observe, {label}", lambda i: _cabi.check(
        lib.mapf_step_observe(env._h, C.c_void_p(ring[i % 8].data_ptr()), C.byref(s), C.c_void_p(obs.data_ptr()),
                              C.c_void_p(vec.data_ptr()), stream())))
    timeit(f"step only, {label}", lambda i: _cabi.check(
        lib.mapf_step(env._h, C.c_void_p(ring[i % 8].data_ptr()), C.byref(s), stream())))
timeit("observe only", lambda i: _cabi.check(lib.mapf_observe(env._h, C.c_void_p(obs.data_ptr()), C.c_void_p(vec.data_ptr()), stream())))
s = so(ALL)


def two(i):
    _cabi.check(lib.mapf_step(env._h, C.c_void_p(ring[i % 8].data_ptr()), C.byref(s), stream()))
    _cabi.check(lib.mapf_observe(env._h, C.c_void_p(obs.data_ptr()), C.c_void_p(vec.data_ptr()), stream()))


timeit("step + observe (two launches)", two)
