#!/usr/bin/env bash
# GAT: fusion on the batch rows, bias gradients from the rows that can be non-zero
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_reference_fixtures.py tests/test_gpu_scale.py -q -p no:cacheprovider --tb=short -x \
   -k "gat or demand or needed or reads or graph_step or full_size or fixtures or forward_predict" ) > $O/r3a_pytest.log 2>&1
tail -3 $O/r3a_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --model gat --steps 20 --warmup 5 --no-cpu-baseline > $O/r3a_bench_gat.json 2> $O/r3a_bench_gat.err
python -c "import json; d=json.loads(open('$O/r3a_bench_gat.json').read().strip().splitlines()[-1]); print('gat', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'))"
