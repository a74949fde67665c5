#!/bin/bash
# A/B experiments: build libdbg_b200 with extra -D flags for filter.cu only (count / partition kernels) into
# rust_debruijn_b200/variants/libdbg_<name>.so; select it at run time with DBG_B200_LIB=<path>.
# usage: tools/build_variant.sh <name> "<extra nvcc flags>" [file.cu]
set -e
name=$1; extra=$2; src=${3:-filter.cu}
cd "$(dirname "$0")/../rust_debruijn_b200/csrc"
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a $extra -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v -c $src -o ../variants/${src%.cu}_$name.o 2> ../variants/${src%.cu}_$name.log
objs=""
for f in scan_sort filter compress graph_ops shard_compress multi capi; do
  if [ "$f.cu" = "$src" ]; then objs="$objs ../variants/${f}_$name.o"; else objs="$objs $f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libdbg_$name.so $objs -lcudart_static -ldl -lrt -lpthread
grep -A1 "count_kernel\|msp_tile_kernel" ../variants/${src%.cu}_$name.log | grep -E "registers|spill" | head -8
echo built variants/libdbg_$name.so
