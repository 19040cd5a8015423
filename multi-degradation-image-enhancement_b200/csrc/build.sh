#!/usr/bin/env bash
# Build libcdan_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
SRCS="plan.cu ops.cu conv_simt.cu conv_umma.cu conv_stream.cu dense_fused.cu cbam.cu glue.cu postproc.cu band.cu"
mkdir -p build
OBJS=""
PIDS=""
for s in $SRCS; do
  o=build/${s%.cu}.o
  # rebuild when the source, any header next to it, or the public header is newer than the object
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.hpp' \) -newer "$o")" ] || [ ../../include/cdan_b200.h -nt "$o" ]; then
    rm -f "$o"  # a failed compile must not leave a stale object for the link step
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$s" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for pid in $PIDS; do
  wait "$pid" || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libcdan_b200.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/libcdan_b200.so"
