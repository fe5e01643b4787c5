#!/usr/bin/env bash
# Round 2, GPU call (one B200): full test suite, bench with the multi-branch demand-driven step, ncu launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2d_pytest.log 2>&1
tail -12 $O/r2d_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2d_bench_gcn.json 2> $O/r2d_bench_gcn.err; tail -c 300 $O/r2d_bench_gcn.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-graph > $O/r2d_bench_gcn_eager.json 2> $O/r2d_bench_gcn_eager.err
timeout 300 python bench.py --steps 10 --warmup 3 --model sage --no-cpu-baseline > $O/r2d_bench_sage.json 2> $O/r2d_bench_sage.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload yelp --no-cpu-baseline > $O/r2d_bench_yelp_gcn.json 2> $O/r2d_bench_yelp_gcn.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ml-small --batch 1024 --no-cpu-baseline > $O/r2d_bench_small_gcn.json 2> $O/r2d_bench_small_gcn.err
# ncu launch list of the bench command's steady state (graph construction and warm-up skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 1500 --csv --log-file $O/r2d_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.2 > $O/r2d_ncu_list.log 2>&1
echo done
