#!/usr/bin/env bash
# N GPUs: the NCCL sharding tests at world = N and the bench line
set -u
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -q -p no:cacheprovider --tb=short -k "nccl and -$N" -s ) > $O/multi_pytest_n$N.log 2>&1
grep -E "passed|failed|vs reference" $O/multi_pytest_n$N.log | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > $O/multi_bench_n$N.json 2> $O/multi_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/multi_bench_n$N.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('step_breakdown_ms'), d['loss'], d.get('strong',{}).get('ms_per_step'), d['host_enqueue_ms_per_step'])
PY
