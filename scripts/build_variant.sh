#!/bin/bash
# Build a library variant in which ONE source is compiled with extra flags (run here; the .so travels to the GPU box).
# usage: scripts/build_variant.sh <name> <source basename without .cu> [-DMACRO=v ...]   → build/variants/libobm_<name>.so
set -e
name=$1; base=$2; shift 2
mkdir -p build/variants build/obj build/vobj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c oceanbiome.jl_b200/csrc/$base.cu -o build/vobj/${base}_$name.o
objs=$(ls build/obj/*.o | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libobm_$name.so $objs build/vobj/${base}_$name.o
echo built build/variants/libobm_$name.so
