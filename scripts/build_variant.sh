#!/bin/bash
# Build a variant of libipc_b200.so with extra compiler flags into ipc_b200/libipc_b200_<name>.so (A/B runs: scripts/ab_libs.py).
# Usage: scripts/build_variant.sh <name> [extra nvcc flags, e.g. -DIPC_RING_ODOM_SHARED=0]
set -e
name=$1; shift
cd "$(dirname "$0")/../ipc_b200/csrc"
mkdir -p build_$name
for f in ipc_capi launch_se2 launch_se3 comm; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o build_$name/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libipc_b200_$name.so build_$name/*.o -ldl
echo "built ipc_b200/libipc_b200_$name.so"
