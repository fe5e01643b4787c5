#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2r_pytest.log 2>&1
grep -E "passed|failed" $O/r2r_pytest.log | tail -2
timeout 400 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2r_bench_n$N.json 2> $O/r2r_bench_n$N.err
echo "rc=$?"; tail -c 300 $O/r2r_bench_n$N.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2r_bench_n1.json 2> $O/r2r_bench_n1.err
echo done
