#!/usr/bin/env bash
# Round 2, multi-GPU call (N = $1 GPUs of one box): real-NCCL sharding parity, bench eager vs captured step, eval.
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 1200 python -m pytest tests/test_gpu_distributed.py -q -p no:cacheprovider --tb=short -k "nccl or eight_way" ) > $O/r2c_pytest_n$N.log 2>&1
tail -6 $O/r2c_pytest_n$N.log
timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2c_bench_n$N.json 2> $O/r2c_bench_n$N.err
echo "eager rc=$?"; tail -c 300 $O/r2c_bench_n$N.err
PEAGNN_BENCH_GRAPH_MULTI=1 timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2c_bench_n${N}_graph.json 2> $O/r2c_bench_n${N}_graph.err
echo "graph rc=$?"; tail -c 600 $O/r2c_bench_n${N}_graph.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N --phase eval --steps 5 --warmup 2 > $O/r2c_bench_eval_n$N.json 2> $O/r2c_bench_eval_n$N.err
echo "eval rc=$?"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8

timeout 600 python tools/shard_diag.py > $O/r2c_shard_diag.txt 2>&1; tail -45 $O/r2c_shard_diag.txt
echo done
