#!/usr/bin/env bash
# GAT backward (batch-dot destination pass, interleaved per-edge pairs) + needed-rows first steps: tests, then the bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_reference_fixtures.py -q -p no:cacheprovider --tb=short -x \
    -k "gat or conv_forward or demand or needed or reads" ) > $O/r2s_pytest.log 2>&1
tail -5 $O/r2s_pytest.log
PEAGNN_BENCH_DUMP_SPMM=1 timeout 300 python bench.py --model gat --steps 20 --warmup 5 --no-cpu-baseline > $O/r2s_bench_gat.json 2> $O/r2s_bench_gat.err
echo "rc=$?"; grep -c PROFILE $O/r2s_bench_gat.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_gat.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('step_breakdown_ms'))
PY
