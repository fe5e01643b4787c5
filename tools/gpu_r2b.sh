#!/usr/bin/env bash
# Round 2, GPU call B (one B200): full test suite, bench lines with the demand-driven step, dense modes.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2b_pytest.log 2>&1
tail -15 $O/r2b_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2b_bench_gcn.json 2> $O/r2b_bench_gcn.err; tail -c 300 $O/r2b_bench_gcn.err
PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --steps 10 --warmup 3 --model gat --no-cpu-baseline > $O/r2b_bench_gat.json 2> $O/r2b_bench_gat.err
timeout 300 python bench.py --steps 10 --warmup 3 --model sage --no-cpu-baseline > $O/r2b_bench_sage.json 2> $O/r2b_bench_sage.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload yelp --model gat --no-cpu-baseline > $O/r2b_bench_yelp_gat.json 2> $O/r2b_bench_yelp_gat.err
for mode in default umma ts; do PEAGNN_DENSE=$mode timeout 120 python tools/pipe_quick.py; done > $O/r2b_dense_modes.txt 2>&1
( nvcc -O3 -I graph_recsys_benchmark_b200/csrc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_rate tools/umma_rate.cu && timeout 60 /tmp/umma_rate ) > $O/r2b_umma_rate.txt 2>&1
echo done
