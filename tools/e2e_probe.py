#!/usr/bin/env python
"""Where the host-buffer call spends its wall time: the whole loop, the C call alone, and the device timeline
(timed with CUDA events by the caller).  One process, 65 536 worlds 40x40x32."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from primal_ppo_b200 import BatchedMapfGym, random_scenario
    from primal_ppo_b200.build import build
    build()
    dev = torch.device("cuda", 0)
    W, H, N = 65536, 40, 32
    sc = random_scenario(W, H, H, N, density=(0.0, 0.3), queue_len=8, seed=3, fov=9, unique_maps=64)
    env = BatchedMapfGym(sc, device=dev, use_tape=False)
    obs = torch.empty((W, N, 6, 9, 9), device=dev); vec = torch.empty((W, N, 4), device=dev)
    hb = env.make_host_buffers(with_obs=False)
    ring = [torch.randint(0, 5, (W, N), dtype=torch.int8).pin_memory() for _ in range(4)]
    for i in range(5):
        env.step_observe_host(hb, obs, vec, actions=ring[i % 4])
    K = 30
    stage = os.environ.get("PROBE_STAGE", "")
    if "clk" in stage:                       # what bench.py does before its e2e loop: NVML sampler thread, then closed
        from bench import ClockSampler
        clk = ClockSampler(0).start()
        with clk:
            for i in range(50):
                env.step_observe(ring[i % 4].cuda(), obs_out=(obs, vec))
            torch.cuda.synchronize(dev)
        print("sampler closed, thread alive:", clk.thread.is_alive() if clk.thread else None, flush=True)
    if "fused" in stage:
        dring = [r.cuda() for r in ring]
        for i in range(50):
            env.step_observe(dring[i % 4], obs_out=(obs, vec))
        for i in range(20):
            env.step(dring[i % 4]); env.getAllObservations(out=(obs, vec))
        torch.cuda.synchronize(dev)
    for label, read in (("call only", False), ("call + host read of reward[0,0]", True)):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter(); inner = 0.0
        for i in range(K):
            a = time.perf_counter()
            env.step_observe_host(hb, obs, vec, actions=ring[i % 4])
            inner += time.perf_counter() - a
            if read:
                _ = float(hb["reward"][0, 0])
        dt = time.perf_counter() - t0
        print(f"{label}: {dt / K * 1e3:.3f} ms per step, inside step_observe_host {inner / K * 1e3:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
