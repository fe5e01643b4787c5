#!/usr/bin/env bash
# Round 2, GPU call A (one B200): tests, the bench lines, probes, ncu launch list + one full capture.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2a_env.txt 2>&1
free -g >> $O/r2a_env.txt; nproc >> $O/r2a_env.txt
( time timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $O/r2a_pytest.log 2>&1
tail -5 $O/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2a_bench_gcn.json 2> $O/r2a_bench_gcn.err; tail -c 400 $O/r2a_bench_gcn.err
PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --steps 20 --warmup 5 --full-propagation --no-cpu-baseline > $O/r2a_bench_gcn_full.json 2> $O/r2a_bench_gcn_full.err
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --no-cpu-baseline > $O/r2a_bench_gat.json 2> $O/r2a_bench_gat.err
timeout 300 python bench.py --steps 10 --warmup 3 --model sage --no-cpu-baseline > $O/r2a_bench_sage.json 2> $O/r2a_bench_sage.err
timeout 400 python bench.py --phase eval --steps 5 --warmup 2 > $O/r2a_bench_eval.json 2> $O/r2a_bench_eval.err
( /usr/bin/time -v timeout 500 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err
grep -E "Maximum resident|Elapsed" $O/r2a_bench_ref.err
timeout 300 python tools/l2_gather_probe.py > $O/r2a_l2_probe.jsonl 2>&1
( PEAGNN_DENSE=ts timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k linear_direct -p no:cacheprovider ) > $O/r2a_ts_pytest.log 2>&1; tail -3 $O/r2a_ts_pytest.log
for mode in default ts pipe; do PEAGNN_DENSE=$mode timeout 120 python tools/pipe_quick.py; done > $O/r2a_dense_modes.txt 2>&1
( nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_rate tools/umma_rate.cu && timeout 60 /tmp/umma_rate ) > $O/r2a_umma_rate.txt 2>&1
# ncu: launch list of the bench command (cold-cache, serialised: shares only), then one full capture of the aggregation
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2a_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph --prewarm 0.2 > $O/r2a_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_rows_kernel|csr_chunk_kernel" -c 8 \
  -o $O/r2a_spmm python tools/spmm_microbench.py --iters 1 --widths 64 > $O/r2a_ncu_spmm.log 2>&1
echo done
