#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, then the ncu passes of B200_PROFILING.md (each only after its
# command exited 0 without ncu).  Outputs under gpurun_out/.
set -u
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/${TAG}_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/${TAG}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:'observe_kernel|step_kernel' -s 6 -c 9 -f \
      -o gpurun_out/${TAG}_full python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/${TAG}_ncu2.log 2>&1
  echo "ncu full rc=$?"
fi
