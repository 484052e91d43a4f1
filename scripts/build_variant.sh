#!/bin/bash
# Build a variant of the library with extra -D flags for dense_tc.cu:  scripts/build_variant.sh <name> <flags...>
# -> variants/libbk_<name>.so (git-ignored; travels to the GPU box).  Run with:  BK_LIB=variants/libbk_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
nvcc $F "$@" -c bayes-kit_b200/csrc/dense_tc.cu -o variants/dense_tc_$NAME.o
OBJS=$(ls bayes-kit_b200/build/*.o | grep -v dense_tc.o)
nvcc -shared -o variants/libbk_$NAME.so $OBJS variants/dense_tc_$NAME.o -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC
rm variants/dense_tc_$NAME.o
echo variants/libbk_$NAME.so
