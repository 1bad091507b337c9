import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200.ppo import ScrimpPolicy
R = 32768
pol = ScrimpPolicy().cuda().eval().use_channels_last()
obs = (torch.rand(R, 6, 9, 9, device="cuda") < 0.15).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
e = pol.enc
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def cr(x, m):
    return torch.cudnn_convolution_relu(x, m.weight.to(torch.bfloat16), m.bias.to(torch.bfloat16), m.stride, m.padding, m.dilation, 1)
def plain():
    x = F.relu(e["c1"](obs)); x = F.relu(e["c1a"](x)); x = F.relu(e["c1b"](x)); x = F.max_pool2d(x, 2)
    x = F.relu(e["c2"](x)); x = F.relu(e["c2a"](x)); x = F.relu(e["c2b"](x)); x = F.max_pool2d(x, 2)
    return F.relu(e["c3"](x).flatten(1))
def fused():
    x = cr(obs, e["c1"]); x = cr(x, e["c1a"]); x = cr(x, e["c1b"]); x = F.max_pool2d(x, 2)
    x = cr(x, e["c2"]); x = cr(x, e["c2a"]); x = cr(x, e["c2b"]); x = F.max_pool2d(x, 2)
    return cr(x, e["c3"]).flatten(1)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    a = plain(); b = fused()
    print("max diff", float((a.float() - b.float()).abs().max()), a.dtype, b.dtype)
    print("plain", t(plain), "ms; fused conv+bias+relu", t(fused), "ms")
