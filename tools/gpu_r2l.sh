#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > $O/r2l_pytest.log 2>&1
grep -E "passed|failed" $O/r2l_pytest.log | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --no-cpu-baseline > $O/r2l_bench_gat.json 2> $O/r2l_bench_gat.err
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --workload yelp --no-cpu-baseline > $O/r2l_bench_yelp_gat.json 2> $O/r2l_bench_yelp_gat.err
timeout 300 python bench.py --steps 10 --warmup 3 --model gat --workload ml-small --batch 1024 --no-cpu-baseline > $O/r2l_bench_small_gat.json 2> $O/r2l_bench_small_gat.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.txt 2>&1; tail -2 $O/r2l_smoke.txt
echo done
