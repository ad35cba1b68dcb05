#!/bin/bash
# A/B builds of the library: tools/build_variant.sh NAME "src1.cu src2.cu" -DFLAG ...
# compiles the listed sources with the extra flags, links them with the default objects of the other sources into
# gdb_nerf_b200/variants/lib_NAME.so (git-ignored, travels to the GPU box; use with tools/bench_k3.py --lib)
set -e
cd "$(dirname "$0")/.."
name=$1; srcs=$2; shift 2
mkdir -p gdb_nerf_b200/variants gdb_nerf_b200/build/var_$name
objs=""
for o in gdb_nerf_b200/build/*.o; do
  b=$(basename $o .o)
  if echo " $srcs " | grep -q " $b.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c gdb_nerf_b200/csrc/$b.cu -o gdb_nerf_b200/build/var_$name/$b.o &
    objs="$objs gdb_nerf_b200/build/var_$name/$b.o"
  else
    objs="$objs $o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gdb_nerf_b200/variants/lib_$name.so $objs -cudart static
ls -la gdb_nerf_b200/variants/lib_$name.so
