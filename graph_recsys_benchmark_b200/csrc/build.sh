#!/usr/bin/env bash
# Builds libpeagnn_sm100.so in-tree (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libpeagnn_sm100.so"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC,-O3,-Wall -shared \
     "$here/graph.cu" "$here/spmm.cu" "$here/dense.cu" "$here/gat.cu" "$here/fuse.cu" \
     "$here/bpr.cu" "$here/eval.cu" \
     -o "$out" "$@"
echo "built $out"
