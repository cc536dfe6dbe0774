#!/bin/bash
# Build an A/B variant of the library: scripts/build_variant.sh NAME "-DFOO=1 ..."  ->  jet-pbrt_b200/build/variants/NAME/libjetpbrt_b200.so
# (build/ is git-ignored but travels with gpurun; on the box:  cp jet-pbrt_b200/build/variants/NAME/*.so jet-pbrt_b200/ )
set -e
cd "$(dirname "$0")/../jet-pbrt_b200"
name=$1; shift
out=build/variants/$name
mkdir -p $out
make -s build/scene.o build/scenes_builtin.o build/capi_host.o build/render.o build/scene_flatten.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 "$@" -Xptxas -v -c csrc/c_api.cu -o $out/c_api.o 2> $out/ptxas.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libjetpbrt_b200.so build/scene.o build/scenes_builtin.o build/capi_host.o build/render.o build/scene_flatten.o $out/c_api.o -Xcompiler -pthread -Xlinker --exclude-libs -Xlinker ALL
grep -E "k_shadeILi0ELb0|k_extendILb0ELi6" -A2 $out/ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    ://'
