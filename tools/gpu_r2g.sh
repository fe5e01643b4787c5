#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2g_pytest.log 2>&1
tail -8 $O/r2g_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench_gcn.json 2> $O/r2g_bench_gcn.err; tail -c 200 $O/r2g_bench_gcn.err
PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gather-dtype bf16 > $O/r2g_bench_gcn_bf16.json 2> $O/r2g_bench_gcn_bf16.err
PEAGNN_BENCH_NO_PROFILE=1 PEAGNN_BENCH_NO_CLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"GatBwdDstOp|GatBwdSrcOp|GatAggOp" --launch-skip 120 -c 8 -o $O/r2g_gat \
  python bench.py --model gat --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.1 > $O/r2g_ncu_gat.log 2>&1
tail -3 $O/r2g_ncu_gat.log
echo done
