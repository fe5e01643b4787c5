#!/usr/bin/env bash
# final single-GPU records: the driver's own commands
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r2p_bench_default.json 2> $O/r2p_bench_default.err
tail -3 $O/r2p_bench_default.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/r2p_bench_reference.json 2> $O/r2p_bench_reference.err
tail -3 $O/r2p_bench_reference.err
( time python bench.py --phase eval --steps 5 --warmup 2 ) > $O/r2p_bench_eval.json 2> $O/r2p_bench_eval.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1700 -c 1400 --csv --log-file $O/r2p_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.2 > $O/r2p_ncu_list.log 2>&1
echo done
