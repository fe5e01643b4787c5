#!/usr/bin/env bash
# 2 GPUs: the NCCL sharding tests (GCN + GAT against the reference's fp64 run) and the N = 2 bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -q -p no:cacheprovider --tb=short -k "nccl" -s ) > $O/r2x_pytest.log 2>&1
grep -E "passed|failed|vs reference" $O/r2x_pytest.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_n2.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'), d['loss'], d.get('strong',{}).get('ms_per_step'))
PY
