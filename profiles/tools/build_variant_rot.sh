#!/bin/bash
# Variant of the library with extra -D flags for rotate_tma.cu into _exp/<name>.so.   build_variant_rot.sh <name> [-D...]
set -e
name=$1; shift
cd "$(dirname "$0")/../../qio_b200/csrc"
mkdir -p ../../_exp
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
nvcc $F "$@" -c rotate_tma.cu -o ../../_exp/rot_$name.o
objs=$(ls *.o | grep -v '^rotate_tma' | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_exp/$name.so $objs ../../_exp/rot_$name.o -cudart static
rm -f ../../_exp/rot_$name.o
echo built _exp/$name.so
