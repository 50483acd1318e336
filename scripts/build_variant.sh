#!/bin/bash
# Build a library variant whose PISCES tendency kernel is compiled with extra flags (run here; the .so travels to the GPU box).
# usage: scripts/build_variant.sh <name> [-DMACRO=v ...]   → build/variants/libobm_<name>.so
set -e
name=$1; shift
src=${OBM_PISCES_SRC:-oceanbiome.jl_b200/csrc/pisces_tendencies.cu}
mkdir -p build/variants build/obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Ioceanbiome.jl_b200/csrc "$@" -c $src -o build/obj/pisces_$name.o
objs=$(ls build/obj/*.o | grep -v "/pisces_")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libobm_$name.so $objs build/obj/pisces_$name.o
echo built build/variants/libobm_$name.so
