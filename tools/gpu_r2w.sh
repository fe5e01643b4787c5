#!/usr/bin/env bash
# A/B: merged lean step vs head + body nodes at 4 / 6 / 8 branches; ncu capture of the GAT passes
set -u
mkdir -p gpurun_out
O=gpurun_out
for cfg in "1 4" "0 4" "1 6" "1 8" "0 4" "1 4"; do
  set -- $cfg
  PEAGNN_MERGED_STEP=$1 PEAGNN_BRANCHES=$2 PEAGNN_BENCH_NO_PROFILE=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2> $O/r2w_bench.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('merged=$1 branches=$2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" | tee -a $O/r2w_ab.txt
done
PEAGNN_BENCH_NO_PROFILE=1 PEAGNN_BENCH_NO_CLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"GatBwdDstOp|GatBwdSrcOp|GatAggOp" --launch-skip 120 -c 8 -o $O/r2w_gat \
  python bench.py --model gat --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.1 > $O/r2w_ncu_gat.log 2>&1
tail -3 $O/r2w_ncu_gat.log
ls -la $O/r2w_gat.ncu-rep
