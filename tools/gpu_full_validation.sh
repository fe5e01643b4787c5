#!/usr/bin/env bash
# full GPU suite + default bench (with the CPU baseline leg), reference arm, eval phase, PEAGAT line, smoke
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/val_pytest.log 2>&1
grep -E "passed|failed" $O/val_pytest.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > $O/val_smoke.txt 2>&1; tail -1 $O/val_smoke.txt
PEAGNN_BENCH_DUMP_SPMM=1 timeout 600 python bench.py > $O/val_bench_default.json 2> $O/val_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --model gat --steps 20 --warmup 5 --no-cpu-baseline > $O/val_bench_gat.json 2> $O/val_bench_gat.err; echo "gat rc=$?"
timeout 300 python bench.py --phase eval --no-cpu-baseline > $O/val_bench_eval.json 2> $O/val_bench_eval.err; echo "eval rc=$?"
python - <<'PY'
import json
for f in ['default','gat','eval']:
    try:
        d=json.loads(open('gpurun_out/val_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['metric'], round(d['value'],1), round(d['ms_per_step'],3), d.get('e2e',{}).get('value'), d.get('cpu_baseline'))
    except Exception as e:
        print(f,'failed',e)
PY
