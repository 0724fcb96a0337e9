#!/bin/bash
# A/B builds of the library: tools/build_variant.sh <name> [extra nvcc flags]  ->  jtokkit_b200/libjtokkit_b200_<name>.so
# (load it with JTK_LIB=jtokkit_b200/libjtokkit_b200_<name>.so; development aid)
set -e
name=$1; shift
cd "$(dirname "$0")/../jtokkit_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -shared -o ../libjtokkit_b200_$name.so jtk_kernels.cu jtk_capi.cu jtk_tables.cpp jtk_regex.cpp jtk_dfa.cpp -lpthread
