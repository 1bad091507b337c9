set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "wide or config5 or maximum or falls_back or fused_step_observe or goal_sampling or all_good or bf16" > gpurun_out/a2_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/a2_pytest.log
python - <<'PY' > gpurun_out/a2_fov.json 2> gpurun_out/a2_fov.err
import json, torch, bench
torch.cuda.set_device(0)
from primal_ppo_b200.build import build; build()
dev = torch.device("cuda", 0)
print(json.dumps(bench.bench_fov_sweep(dev, 0, 1)))
PY
echo "fov rc=$?"; tail -3 gpurun_out/a2_fov.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/a2_fov.json"))
for r in d["rows"]:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
PY
