#!/usr/bin/env python
"""In-process A/B of the CTA-per-world fused kernel's launch variants (MAPF_DBG_FLAGS bits 24-25) on BASELINE configs[4]:
1 = 8 warps x 3 CTAs/SM, 2 = 4 warps x 6, 3 = 4 warps x 8; 'two' = the two launches (bit 0)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.cuda.set_device(0)
from primal_ppo_b200.build import build  # noqa: E402
build()
dev = torch.device("cuda", 0)
for name, flags in (("two-launch", 1), ("8w x3", 1 << 24), ("4w x6", 2 << 24), ("4w x8", 3 << 24)):
    os.environ["MAPF_DBG_FLAGS"] = str(flags)
    r = bench.bench_fov_sweep(dev, 0, 1)
    print(name, json.dumps([{k: round(x[k], 4) for k in ("fov", "ms_per_step", "step_observe_frac")} for x in r["rows"]]), flush=True)
