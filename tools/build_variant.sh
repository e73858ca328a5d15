#!/bin/bash
# A/B builds of libcramore_cuda.so: tools/build_variant.sh NAME "-DMACRO=VALUE ..." -> cramore_b200/variants/libccu_NAME.so
# (same ABI; select it with CCU_LIB=cramore_b200/variants/libccu_NAME.so).  Used to measure kernel alternatives in
# one GPU session; the shipped library is always the default build of cramore_b200/csrc/Makefile.
set -e
name=$1; shift
flags="$*"
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/cramore_b200/csrc
obj=$(mktemp -d)
mkdir -p $root/cramore_b200/variants
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$root/include -I$src $flags"
for f in ccu_api ccu_fmx_api demux_kernels fmx_kernels fmx_init; do $NV -c $src/$f.cu -o $obj/$f.o & done
for f in demux_host demux_refine fmx_host; do g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I$root/include -c $src/$f.cpp -o $obj/$f.o & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/cramore_b200/variants/libccu_$name.so $obj/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $obj
echo built cramore_b200/variants/libccu_$name.so
