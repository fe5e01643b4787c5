#!/usr/bin/env bash
# merged lean step: model-level tests + the full-size reproducibility test, then the bench line (merged / two-node)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_scale.py tests/test_gpu_reference_fixtures.py tests/test_gpu_bf16.py -q -p no:cacheprovider --tb=short -x ) > $O/r2v_pytest.log 2>&1
tail -4 $O/r2v_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2v_bench.json 2> $O/r2v_bench.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'), d['loss'])
PY
