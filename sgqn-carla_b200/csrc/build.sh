#!/bin/bash
# Builds libsgqn_b200.so in-tree (travels to the GPU box with the gpurun snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in abi dense replay saliency heads optim conv_tc conv_chain conv_tcg conv1_tc gemm_tc p2p $EXTRA_SRCS; do
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
objs=""; for f in abi dense replay saliency heads optim conv_tc conv_chain conv_tcg conv1_tc gemm_tc p2p $EXTRA_SRCS; do objs="$objs build/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libsgqn_b200.so $objs -lcudart
echo "built $(ls -la ../libsgqn_b200.so)"
