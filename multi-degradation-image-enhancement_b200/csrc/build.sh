#!/usr/bin/env bash
# Build libcdan_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
SRCS="plan.cu ops.cu conv_simt.cu conv_umma.cu conv_stream.cu cbam.cu glue.cu postproc.cu"
mkdir -p build
OBJS=""
for s in $SRCS; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.hpp' \) -newer "$o")" ] || [ ../../include/cdan_b200.h -nt "$o" ]; then
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$s" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libcdan_b200.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/libcdan_b200.so"
