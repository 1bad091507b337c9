#!/usr/bin/env python
"""Run bench.py under several MAPF_DBG_FLAGS values and print one summary line each (kernel experiments)."""
import json
import os
import subprocess
import sys

flags = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"]
extra = sys.argv[2:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f in flags:
    env = dict(os.environ, MAPF_DBG_FLAGS=f)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-extras", "--cpu-budget", "0"] + extra,
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        k = d["kernels"]
        print(f"flags={f} value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} obs_ms={k['observe_kernel']['ms']:.4f} "
              f"step_ms={k['step_kernel']['ms']:.4f} obs_frac={k['observe_kernel']['frac']:.3f} e2e={d['e2e']['value']:.4g}",
              flush=True)
    except Exception as ex:
        print(f"flags={f} FAILED {ex}: {r.stderr[-400:]}", flush=True)
