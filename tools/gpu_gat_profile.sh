#!/usr/bin/env bash
# 1 GPU: GAT pass timings, one ncu capture of the GAT passes on user2item (F = 64), other workloads' bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gat_microbench.py --iters 10 --feat 64 > $O/gat_gat_microbench.txt 2>&1
timeout 300 python tools/gat_microbench.py --iters 10 --feat 16 >> $O/gat_gat_microbench.txt 2>&1
cat $O/gat_gat_microbench.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"Gat|RowMax" -c 16 -o $O/gat_gat python tools/gat_microbench.py --iters 0 --feat 64 > $O/gat_ncu_gat.log 2>&1
tail -2 $O/gat_ncu_gat.log
ncu -i $O/gat_gat.ncu-rep --page raw --csv > $O/gat_gat_raw.csv 2>/dev/null; ls -la $O/gat_gat.ncu-rep $O/gat_gat_raw.csv
for w in "yelp gat" "yelp gcn" "ml-small gat" "ml-small gcn"; do
  set -- $w
  b=4096; [ "$1" = "ml-small" ] && b=1024
  timeout 300 python bench.py --workload $1 --model $2 --batch $b --steps 20 --warmup 5 --no-cpu-baseline 2> $O/gat_bench.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" | tee -a $O/gat_workloads.txt
done
