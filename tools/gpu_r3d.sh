#!/usr/bin/env bash
set -u
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_scale.py -q -p no:cacheprovider --tb=short -x -k "demand or full_size or graph_step" ) > $O/r3d_pytest.log 2>&1
tail -2 $O/r3d_pytest.log
for g in 1 0 1; do
  PEAGNN_GROUPED=$g PEAGNN_BENCH_NO_PROFILE=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2> $O/r3d.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ml-25m grouped=$g', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" | tee -a $O/r3d_ab.txt
done
