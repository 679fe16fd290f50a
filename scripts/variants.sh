#!/bin/bash
# Kernel tuning experiments: build one library per set of -D flags into build_variants/ (ignored by git, shipped to the
# GPU box by gpurun), to be compared with LORO_B200_LIB=... python bench.py on the box.
# usage: scripts/variants.sh name "flags" [name "flags" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (cd loro_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC -shared -diag-suppress 550 $flags -o ../../build_variants/$name.so engine.cu) &
done
wait
ls -la build_variants
