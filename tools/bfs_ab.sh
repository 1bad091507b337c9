#!/bin/bash
# bfs_kernel A/B on one box: GPU parity tests, then the bench's BFS extra under MAPF_DBG_FLAGS variants
# (65536 = row-word kernel, 131072/262144/393216 = cell-string kernel forced to 8/16/32 lanes per map).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for f in 0 65536 131072 262144 393216 0 65536; do
  MAPF_DBG_FLAGS=$f python bench.py --steps 5 --warmup 3 --cpu-budget 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags', $f, 'bfs', d.get('bfs'))"
done
