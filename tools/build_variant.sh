#!/bin/bash
# Kernel experiments: build libbgs_b200.so with extra -D flags into build/variants/<name>/ (git-ignored, travels to the
# GPU box); time it with `python tools/time_rollout.py --lib build/variants/<name>/libbgs_b200.so`.
#   tools/build_variant.sh <name> [-DFLAG ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=${BGS_SRC:-$root/board-game-simulator-python_b200/csrc}
out=$root/build/variants/$name
mkdir -p "$out"
for f in api connect bounce keys; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c -o "$out/$f.o" "$src/$f.cu" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o "$out/libbgs_b200.so" "$out"/*.o
rm -f "$out"/*.o
echo "$out/libbgs_b200.so"
