#!/usr/bin/env python
"""Per-call step_kernel durations over consecutive steps (CUDA events), to find steps where the kernel is much slower
than its median.  Prints the outliers with the step index and a few world statistics of that step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import build_scenario, N_AGENTS
    from primal_ppo_b200 import BatchedMapfGym
    from primal_ppo_b200.build import build
    build()
    dev = torch.device("cuda", 0)
    W, N = 65536, N_AGENTS
    sc = build_scenario(W, seed=100)
    env = BatchedMapfGym(sc, device=dev, seed=1234, use_tape=False)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    ring = [torch.randint(0, 5, (W, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(T)]
    stats = []
    for t in range(T):
        ev[t][0].record()
        out = env.step(ring[t % 8])
        ev[t][1].record()
        stats.append((int(out.goals_reached.sum()), int(out.violated.sum()), int((out.status != 0).sum())))
    torch.cuda.synchronize(dev)
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    med = float(np.median(ms))
    print(f"median {med:.4f} ms, mean {ms.mean():.4f}, max {ms.max():.4f}")
    for t in np.nonzero(ms > 1.5 * med)[0]:
        print(f"step {t}: {ms[t]:.3f} ms  ring {t % 8}  goals {stats[t][0]} violated {stats[t][1]} status!=0 {stats[t][2]}")
    print("first 24:", [round(float(x), 3) for x in ms[:24]])


if __name__ == "__main__":
    main()
