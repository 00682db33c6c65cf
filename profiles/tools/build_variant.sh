#!/bin/bash
# Builds a variant of the library with extra -D flags for sweep3.cu into _exp/<name>.so (A/B runs through QIO_B200_LIB).
#   profiles/tools/build_variant.sh <name> [-DS3_...=v ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../../qio_b200/csrc"
mkdir -p ../../_exp/obj_$name
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
nvcc $F "$@" -c sweep3.cu -o ../../_exp/obj_$name/sweep3.o &
nvcc $F "$@" -c sweep3_timed.cu -o ../../_exp/obj_$name/sweep3_timed.o &
wait
objs=$(ls *.o | grep -v '^sweep3' | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_exp/$name.so $objs ../../_exp/obj_$name/sweep3.o ../../_exp/obj_$name/sweep3_timed.o -cudart static
rm -rf ../../_exp/obj_$name
echo built _exp/$name.so
