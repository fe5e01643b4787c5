#!/usr/bin/env bash
# Round 2, 8-GPU call: real-NCCL sharding parity at 8 ranks, bench with the captured step, eval.
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -q -p no:cacheprovider --tb=short -k "nccl and gcn" ) > $O/r2k_pytest_n$N.log 2>&1
tail -5 $O/r2k_pytest_n$N.log
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2k_bench_n$N.json 2> $O/r2k_bench_n$N.err
echo "graph rc=$?"; tail -c 400 $O/r2k_bench_n$N.err
PEAGNN_BENCH_EAGER_MULTI=1 timeout 400 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-strong > $O/r2k_bench_n${N}_eager.json 2> $O/r2k_bench_n${N}_eager.err
echo "eager rc=$?"
timeout 300 $TR --master-port 29523 bench.py --gpus $N --phase eval --steps 5 --warmup 2 > $O/r2k_bench_eval_n$N.json 2> $O/r2k_bench_eval_n$N.err
echo "eval rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r2k_bench_n4.json 2> $O/r2k_bench_n4.err
echo "n4 rc=$?"
echo done
