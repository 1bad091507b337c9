#!/usr/bin/env python
"""BASELINE.json configs[4]: 80x80 worlds, 128 agents, large-FOV observation sweep (memory-bound roofline stress).
One JSON line per FOV: step (step_wide_kernel) and observe (observe_kernel) timed with CUDA events, algorithmic bytes
per SURVEY.md §8d (B = 62 + 4*C*F^2 + (H*W+8)/N per agent-step), fraction of the measured HBM peak.
Weak scaling over GPUs is trivial (worlds are independent); run under torchrun to see it."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import measured_peaks  # noqa: E402


def main():
    from primal_ppo_b200 import BatchedMapfGym, random_scenario
    from primal_ppo_b200.build import build
    build()
    rank = int(os.environ.get("RANK", "0")); ws = int(os.environ.get("WORLD_SIZE", "1"))
    lrk = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrk)
    dev = torch.device("cuda", lrk)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, N, C = 80, 128, 6
    peak, src = measured_peaks()
    for F, W in ((9, 16384), (15, 8192), (21, 4096), (31, 2048)):
        sc = random_scenario(W, H, H, N, density=(0.0, 0.3), queue_len=8, seed=900 + F + rank, fov=F, unique_maps=64)
        env = BatchedMapfGym(sc, device=dev, use_tape=False)
        obs = torch.empty((W, N, C, F, F), device=dev); vec = torch.empty((W, N, 4), device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(F)
        ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(4)]
        for i in range(3):
            env.step(ring[i % 4]); env.getAllObservations(out=(obs, vec))
        K = 10
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        for i in range(K):
            ev[i][0].record(); env.step(ring[i % 4]); ev[i][1].record(); env.getAllObservations(out=(obs, vec)); ev[i][2].record()
        torch.cuda.synchronize(dev)
        step_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
        obs_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
        t = torch.tensor([step_ms + obs_ms], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shared = (H * H + 8) / N
        b_obs = 16 + 4 * C * F * F + shared + 8
        b_all = 62 + 4 * C * F * F + shared
        if rank == 0:
            print(json.dumps({"workload": f"{W} worlds/GPU 80x80, 128 agents, FOV {F}", "n_gpus": ws, "fov": F,
                              "agent_steps_per_s": W * N * ws / (float(t) * 1e-3), "step_ms": step_ms, "observe_ms": obs_ms,
                              "obs_bytes_per_launch": b_obs * W * N, "observe_gbs": b_obs * W * N / (obs_ms * 1e-3) / 1e9,
                              "observe_frac": b_obs * W * N / (obs_ms * 1e-3) / 1e9 / peak,
                              "step_observe_gbs": b_all * W * N / ((step_ms + obs_ms) * 1e-3) / 1e9,
                              "step_observe_frac": b_all * W * N / ((step_ms + obs_ms) * 1e-3) / 1e9 / peak,
                              "peak_gbs": peak, "peak_source": src,
                              "worlds_with_error_flags": float((env.state()["err"] != 0).float().mean())}), flush=True)
        del env, obs, vec, ring
        torch.cuda.empty_cache()
    if ws > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
