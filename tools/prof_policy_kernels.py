#!/usr/bin/env python
"""Top CUDA kernels of one ScrimpPolicy forward (bf16 autocast, channels_last), torch.profiler."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200.ppo import ScrimpPolicy
from torch.profiler import ProfilerActivity, profile
R = 32768
torch.manual_seed(0)
pol = ScrimpPolicy().cuda().eval().use_channels_last()
obs = (torch.rand(R, 6, 9, 9, device="cuda") < 0.15).float(); vec = torch.randn(R, 4, device="cuda")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        pol.features(obs, vec)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pol.features(obs, vec)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
