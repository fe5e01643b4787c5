#!/usr/bin/env bash
# Builds libpeagnn_sm100.so in-tree (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
# Translation units are compiled side by side into csrc/_build/ and linked at the end.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libpeagnn_sm100.so"
obj="$here/_build"
mkdir -p "$obj"
pids=()
for unit in graph spmm dense gat fuse bpr eval sample probe; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
       -Xcompiler -fPIC,-O3,-Wall -c "$here/$unit.cu" -o "$obj/$unit.o" "$@" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared \
     "$obj"/graph.o "$obj"/spmm.o "$obj"/dense.o "$obj"/gat.o "$obj"/fuse.o "$obj"/bpr.o "$obj"/eval.o "$obj"/sample.o "$obj"/probe.o \
     -o "$out"
echo "built $out"
