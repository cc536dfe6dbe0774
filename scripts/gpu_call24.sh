#!/bin/bash
# round 2, GPU call 24: node stack position as a pointer (predicated push, no index arithmetic), tfar not widened -- timing and parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "sp nw spnw" --scenes=bunny,cornell,glossy,large --spp=48 > gpurun_out/c24_ab_stackptr.log 2>&1
cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so
cp jet-pbrt_b200/build/variants/spnw/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_guards.py tests/test_gpu_bvh_build.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c24_tests_spnw.log 2>&1
echo "tests rc=$?" >> gpurun_out/c24_tests_spnw.log
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
cat gpurun_out/c24_ab_stackptr.log; tail -n 5 gpurun_out/c24_tests_spnw.log
