#!/bin/bash
# Build libscsfm.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
OUT=scsfm/libscsfm.so
SRCS=$(ls csrc/*.cu)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p build
OBJS=""
for s in $SRCS; do
  o=build/$(basename ${s%.cu}).o
  if [ ! -f $o ] || [ $s -nt $o ] || [ -n "$(find csrc include ../include -newer $o -name '*.cuh' -o -newer $o -name '*.h' 2>/dev/null | head -1)" ]; then
    echo "nvcc $s"
    $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA} -c $s -o $o &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $OBJS -lcudart -lcuda
echo "built $OUT"
