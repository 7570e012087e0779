#!/bin/bash
# Build and run the inner-loop microbenchmark on the GPU box (nvcc is in the image).
set -e
cd "$(dirname "$0")"
mkdir -p _build ../gpurun_out
for u in 1 7; do
  nvcc -DKUNROLL=$u -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
       -o _build/microbench_inner_u$u microbench_inner.cu
  echo "== KUNROLL=$u" | tee -a ../gpurun_out/microbench_inner.txt
  ./_build/microbench_inner_u$u | tee -a ../gpurun_out/microbench_inner.txt
done
