#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (each timed).  Outputs under gpurun_out/<TAG>_*.
set -u
TAG=${1:-chk}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc; free -g | head -2
t0=$(date +%s)
python -m pytest tests -m gpu -q --durations=12 ${PYTEST_ARGS:-} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -25 gpurun_out/${TAG}_pytest.log
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"
tail -2 gpurun_out/${TAG}_smoke.log
t0=$(date +%s)
python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "sustained", "roofline", "clocks", "strong", "fov_sweep", "ppo", "bfs", "gae", "e2e_synchronous_call")}
    print(json.dumps(keep, indent=1)[:6000])
except Exception as ex:
    print("no bench line:", ex)
PY
