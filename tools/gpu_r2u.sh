#!/usr/bin/env bash
# full GPU suite, then the PEAGCN bench line with 4 / 8 / 12 parallel branches and the PEAGAT line
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2u_pytest.log 2>&1
grep -E "passed|failed" $O/r2u_pytest.log | tail -2
for b in 4 8 12; do
  PEAGNN_BRANCHES=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2u_bench_b$b.json 2> $O/r2u_bench_b$b.err
  echo "branches $b rc=$?"
done
timeout 300 python bench.py --model gat --steps 20 --warmup 5 --no-cpu-baseline > $O/r2u_bench_gat.json 2> $O/r2u_bench_gat.err
python - <<'PY'
import json
for f in ['b4','b8','b12','gat']:
    try:
        d=json.loads(open('gpurun_out/r2u_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'))
    except Exception as e:
        print(f, 'failed', e)
PY
