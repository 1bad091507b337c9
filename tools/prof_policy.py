#!/usr/bin/env python
"""Where does the policy forward spend its time?  (plain PyTorch; the only dense contraction of the path)"""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200.ppo import ScrimpPolicy

R = int(os.environ.get("ROWS", 32768))
torch.manual_seed(0)
pol = ScrimpPolicy().cuda().eval()
obs = (torch.rand(R, 6, 9, 9, device="cuda") < 0.15).float()
vec = torch.randn(R, 4, device="cuda")


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def enc(o):
    e = pol.enc
    x = F.relu(e["c1"](o)); x = F.relu(e["c1a"](x)); x = F.relu(e["c1b"](x)); x = F.max_pool2d(x, 2)
    x = F.relu(e["c2"](x)); x = F.relu(e["c2a"](x)); x = F.relu(e["c2b"](x)); x = F.max_pool2d(x, 2)
    return F.relu(e["c3"](x).flatten(1))


for name, dt, cl, bench in (("fp32", None, False, False), ("bf16", torch.bfloat16, False, False), ("bf16+cudnn.benchmark", torch.bfloat16, False, True),
                            ("bf16+channels_last+benchmark", torch.bfloat16, True, True)):
    torch.backends.cudnn.benchmark = bench
    o = obs.contiguous(memory_format=torch.channels_last) if cl else obs
    if cl:
        pol.enc.to(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=dt, enabled=dt is not None):
        full = t(lambda: pol.features(o, vec))
        conv = t(lambda: enc(o))
        seq = torch.randn(R, 17, 512, device="cuda")
        tr = t(lambda: pol.blocks[1](pol.blocks[0](seq, cls_only=False), cls_only=True))
    print(f"{name:32s} rows={R} features {full:7.2f} ms  conv stack {conv:7.2f} ms  transformer {tr:7.2f} ms  "
          f"-> {R * 213e6 / (full * 1e-3) / 1e12:6.1f} TFLOP/s (213 MFLOP/row nominal)", flush=True)
