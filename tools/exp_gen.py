import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primal_ppo_b200 import generate_scenario_device
for kind, H, Wd in (("density", 40, 40), ("warehouse", 40, 60)):
    generate_scenario_device(256, H, Wd, 32, kind=kind, queue_len=16, seed=1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d = generate_scenario_device(65536, H, Wd, 32, kind=kind, queue_len=16, seed=2); b.record(); torch.cuda.synchronize()
    print(kind, "65536 worlds:", round(a.elapsed_time(b), 2), "ms; flagged", float(((d.gen_err & ~4) != 0).float().mean()), "checksum", int(d.htrace.long().sum()), int(d.hlen.sum()))
