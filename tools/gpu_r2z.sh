#!/usr/bin/env bash
# sparse-filter schedule: tests, then the PEAGCN / PEAGAT bench lines; then tools/gpu_r2y.sh (GAT microbench, ncu, other workloads)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_scale.py -q -p no:cacheprovider --tb=short -x \
   -k "filtered or demand or full_size or needed or reads or graph_step" ) > $O/r2z_pytest.log 2>&1
tail -3 $O/r2z_pytest.log
for m in gcn gat; do
  PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline > $O/r2z_bench_$m.json 2> $O/r2z_bench_$m.err
  python -c "import json; d=json.loads(open('$O/r2z_bench_$m.json').read().strip().splitlines()[-1]); print('$m', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'))"
done
bash tools/gpu_r2y.sh
