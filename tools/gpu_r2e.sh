#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2e_pytest.log 2>&1
tail -8 $O/r2e_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2e_bench_gcn.json 2> $O/r2e_bench_gcn.err; tail -c 300 $O/r2e_bench_gcn.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --full-propagation > $O/r2e_bench_gcn_full.json 2> $O/r2e_bench_gcn_full.err
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --no-cpu-baseline > $O/r2e_bench_gat.json 2> $O/r2e_bench_gat.err
echo done
