#!/bin/bash
# build a kernel-variant copy of the library for A/B timing: profiles/build_variant.sh <name> <extra nvcc flags...>
# -> variants/libfv3_<name>.so (git-ignored, travels to the GPU box); select with FV3_B200_LIB=variants/libfv3_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj_$name
for s in ctx tp2d c_sw d_sw nh pgrad halo dyn_core tracer dyn_post; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c gfdl_atmos_cubed_sphere_b200/csrc/$s.cu -o variants/obj_$name/$s.o &
done
wait
nvcc -shared -o variants/libfv3_$name.so variants/obj_$name/*.o -cudart static -ldl
echo built variants/libfv3_$name.so
