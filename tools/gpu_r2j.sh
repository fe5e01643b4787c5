#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x ) > $O/r2j_pytest.log 2>&1
tail -6 $O/r2j_pytest.log
for br in 4 8; do
PEAGNN_BRANCHES=$br timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_bench_gcn_b$br.json 2> $O/r2j_bench_gcn_b$br.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --workload yelp --no-cpu-baseline > $O/r2j_bench_yelp_gcn.json 2> $O/r2j_bench_yelp_gcn.err
echo done
