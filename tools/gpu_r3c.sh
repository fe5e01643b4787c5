#!/usr/bin/env bash
set -u
O=gpurun_out
for w in "ml-small 1024" "yelp 4096"; do
  set -- $w
  for g in 1 0; do
    PEAGNN_GROUPED=$g PEAGNN_BENCH_NO_PROFILE=1 timeout 300 python bench.py --workload $1 --batch $2 --steps 40 --warmup 5 --no-cpu-baseline 2> $O/r3c.err | \
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 grouped=$g', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'])" | tee -a $O/r3c_ab.txt
  done
done
PEAGNN_GROUPED=1 PEAGNN_BENCH_NO_PROFILE=1 timeout 300 python bench.py --full-propagation --steps 20 --warmup 5 --no-cpu-baseline 2>> $O/r3c.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full grouped=1', round(d['ms_per_step'],3))" | tee -a $O/r3c_ab.txt
PEAGNN_GROUPED=0 PEAGNN_BENCH_NO_PROFILE=1 timeout 300 python bench.py --full-propagation --steps 20 --warmup 5 --no-cpu-baseline 2>> $O/r3c.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full grouped=0', round(d['ms_per_step'],3))" | tee -a $O/r3c_ab.txt
