#!/usr/bin/env python
"""bfs_kernel timings over map shapes, CUDA events, one JSON line per (shape, kernel variant).
MAPF_DBG_FLAGS is read when the environment is created: 0 = shipped rule, 65536 = row-word kernel,
131072 / 262144 / 393216 = cell-string kernel forced to 8 / 16 / 32 lanes per map."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from primal_ppo_b200 import BatchedMapfGym, random_scenario
    from primal_ppo_b200.build import build
    build()
    dev = torch.device("cuda", 0)
    shapes = [(16384, 40, 40, 32), (8192, 20, 20, 8), (512, 80, 80, 128), (128, 128, 128, 128), (8192, 32, 64, 16)]
    variants = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "65536"])]
    for (W, H, Wd, N) in shapes:
        sc = random_scenario(W, H, Wd, N, density=(0.0, 0.3), queue_len=4, seed=5, fov=9, unique_maps=64)
        for f in variants:
            os.environ["MAPF_DBG_FLAGS"] = str(f)
            env = BatchedMapfGym(sc, device=dev, use_tape=False)
            out = torch.empty((W * N, H, Wd), dtype=torch.int16, device=dev)
            for _ in range(2):
                env.bfs_maps(out=out)
            torch.cuda.synchronize(dev)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                env.bfs_maps(out=out)
            b.record(); torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / 5
            print(json.dumps({"worlds": W, "H": H, "W": Wd, "agents": N, "maps": W * N, "flags": f, "ms": round(ms, 4),
                              "maps_per_s": W * N / (ms * 1e-3), "depth_max": int(out.max().item()),
                              "written_gbs": W * N * H * Wd * 2 / (ms * 1e-3) / 1e9}), flush=True)
            del env, out


if __name__ == "__main__":
    main()
