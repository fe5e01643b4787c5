#!/usr/bin/env bash
# ncu: launch list of the bench command (shares), one full capture of the grouped projection kernels
set -u
mkdir -p gpurun_out
O=gpurun_out
PEAGNN_BENCH_NO_PROFILE=1 PEAGNN_BENCH_NO_CLOCKS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 1500 --csv --log-file $O/ncu_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.2 > $O/ncu_ncu_list.log 2>&1
python tools/summarize_launches.py $O/ncu_launches.csv > $O/ncu_launches_summary.txt 2>&1; head -30 $O/ncu_launches_summary.txt
PEAGNN_BENCH_NO_PROFILE=1 PEAGNN_BENCH_NO_CLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"_grouped" --launch-skip 60 -c 12 -o $O/ncu_grouped \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.1 > $O/ncu_ncu_grouped.log 2>&1
tail -2 $O/ncu_ncu_grouped.log
ncu -i $O/ncu_grouped.ncu-rep --page raw --csv > $O/ncu_grouped_raw.csv 2>/dev/null; ls -la $O/ncu_grouped.ncu-rep $O/ncu_grouped_raw.csv
