#!/usr/bin/env bash
# grouped projection launches: kernel tests, model tests, then A/B of the PEAGCN bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_scale.py tests/test_gpu_reference_fixtures.py -q -p no:cacheprovider --tb=short \
   -k "grouped or demand or full_size or graph_step or fixtures or linear_direct or wgrad_direct or training_steps" ) > $O/r3b_pytest.log 2>&1
tail -3 $O/r3b_pytest.log
for g in 1 0 1; do
  PEAGNN_GROUPED=$g timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $O/r3b_bench_g$g.json 2> $O/r3b_bench_g$g.err
  python -c "import json; d=json.loads(open('$O/r3b_bench_g$g.json').read().strip().splitlines()[-1]); print('grouped=$g', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'), d.get('roofline_projection',{}).get('frac'), d['gpu_launches'])" | tee -a $O/r3b_ab.txt
done
