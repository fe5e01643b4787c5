#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2i_pytest.log 2>&1
tail -8 $O/r2i_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2i_bench_gcn.json 2> $O/r2i_bench_gcn.err; tail -c 200 $O/r2i_bench_gcn.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ml-small --batch 1024 --no-cpu-baseline > $O/r2i_bench_small_gcn.json 2> $O/r2i_bench_small_gcn.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload yelp --no-cpu-baseline > $O/r2i_bench_yelp_gcn.json 2> $O/r2i_bench_yelp_gcn.err
echo done
