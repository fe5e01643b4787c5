#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2h_pytest.log 2>&1
tail -8 $O/r2h_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2h_bench_gcn.json 2> $O/r2h_bench_gcn.err; tail -c 200 $O/r2h_bench_gcn.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --full-propagation > $O/r2h_bench_gcn_full.json 2> $O/r2h_bench_gcn_full.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gather-dtype bf16 > $O/r2h_bench_gcn_bf16.json 2> $O/r2h_bench_gcn_bf16.err
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --no-cpu-baseline > $O/r2h_bench_gat.json 2> $O/r2h_bench_gat.err
timeout 200 python tools/spmm_microbench.py --iters 10 --widths 64,16,112 > $O/r2h_spmm_micro.txt 2>&1
PEAGNN_BENCH_NO_PROFILE=1 PEAGNN_BENCH_NO_CLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:"GatBwdDstOp|GatBwdSrcOp|GatAggOp" --launch-skip 100 -c 8 -o $O/r2h_gat \
  python bench.py --model gat --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.1 > $O/r2h_ncu_gat.log 2>&1
tail -3 $O/r2h_ncu_gat.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_rows_kernel|csr_chunk_kernel" -c 4 \
  -o $O/r2h_spmm python tools/spmm_microbench.py --iters 1 --widths 64 > $O/r2h_ncu_spmm.log 2>&1
echo done
