set -u
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/fin_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/fin_ncu1.log 2>&1
echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'observe_kernel|step_kernel' -s 6 -c 9 -f -o gpurun_out/fin_full python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > gpurun_out/fin_ncu2.log 2>&1
echo "full rc=$?"
python bench.py --steps 4 --warmup 3 --cpu-budget 0 > gpurun_out/fin_plain_extras.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bfs_kernel' -c 3 -f -o gpurun_out/fin_bfs python bench.py --steps 4 --warmup 3 --cpu-budget 0 > gpurun_out/fin_ncu3.log 2>&1
echo "bfs rc=$?"
ls -la gpurun_out/fin_*
