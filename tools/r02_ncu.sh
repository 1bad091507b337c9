#!/bin/bash
# One gpurun call: write-pattern microbench (no ncu), then the ncu passes of B200_PROFILING.md, each right after the same
# command exited 0 without ncu.  Outputs under gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
./tools/wpb > gpurun_out/r02_wpb.txt 2>&1; echo "wpb rc=$?"; grep -E "^(A warp chunks .cs, (2|4)|D warp.*LUT.*47104|G )" gpurun_out/r02_wpb.txt
CMD="python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
echo "launches rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'observe_kernel|step_kernel' -s 6 -c 9 -f -o gpurun_out/r02_full $CMD > gpurun_out/r02_ncu2.log 2>&1
echo "full rc=$?"
python tools/prof_wide.py > gpurun_out/r02_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'step_observe_wide_kernel|bfs_gray_kernel|gae' -s 2 -c 8 -f -o gpurun_out/r02_wide python tools/prof_wide.py > gpurun_out/r02_ncu3.log 2>&1
echo "wide rc=$?"
ls -la gpurun_out/r02_*
