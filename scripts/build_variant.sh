#!/bin/bash
# Build a library variant in which some sources are compiled with extra flags (run here; the .so travels to the GPU box).
# usage: scripts/build_variant.sh <name> <source basenames, comma separated, without .cu> [-DMACRO=v ...]
#        → build/variants/libobm_<name>.so
set -e
name=$1; bases=$2; shift 2
mkdir -p build/variants build/obj build/vobj
objs=$(ls build/obj/*.o)
extra=""
for base in ${bases//,/ }; do
  per_source=""; [ "$base" = npd_tendencies ] && per_source="-fmad=false"  # as __graft_entry__.PER_SOURCE_FLAGS
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $per_source "$@" -c oceanbiome.jl_b200/csrc/$base.cu -o build/vobj/${base}_$name.o
  objs=$(echo "$objs" | grep -v "/$base.o")
  extra="$extra build/vobj/${base}_$name.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libobm_$name.so $objs $extra
echo built build/variants/libobm_$name.so
