#!/bin/bash
# Builds ninpol_b200/_variants/lib_NAME.so: the library with k2_gls.cu compiled with extra nvcc flags (the other objects
# come from ninpol_b200/_obj, so run `python -m ninpol_b200.build` first).  tools/gpurun/r02_variants.sh times them.
# usage: tools/gpurun/mkvariant.sh NAME [nvcc flags for k2_gls.cu ...]
#   e.g. mkvariant.sh base;  mkvariant.sh u2 -DMF_APPLY_UNROLL=2;  mkvariant.sh hh -DMF_HH_NOINLINE;  mkvariant.sh z -DMF_ZERO_FIRST
set -e
name=$1; shift
cd "$(dirname "$0")/../../ninpol_b200"
mkdir -p _variants /tmp/npb_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++ "$@" \
      -c csrc/k2_gls.cu -o /tmp/npb_variants/k2_gls_$name.o
objs=$(ls _obj/*.o | grep -v "/k2_gls.o")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o _variants/lib_$name.so $objs /tmp/npb_variants/k2_gls_$name.o -lcudart -ldl -lpthread -ccbin /usr/bin/g++
cuobjdump -res-usage _variants/lib_$name.so 2>/dev/null | grep -A1 "k_gls_mfILi12" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' '
echo " $name"
