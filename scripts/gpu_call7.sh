#!/bin/bash
# round 2, GPU call 7: leaf phase spread over the warp (JPB_LEAF_SHARE): correctness (bit-exact hit tests, same-path images) and timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so
cp jet-pbrt_b200/build/variants/leafshare6/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_guards.py -m gpu -x -q) > gpurun_out/c7_tests_leafshare.log 2>&1
echo "tests rc=$?" >> gpurun_out/c7_tests_leafshare.log
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
scripts/ab_variants.sh "leafshare6 leafshare_u1" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c7_ab_leafshare.log 2>&1
scripts/ab_variants.sh "leafshare6" trav_blocks=5 --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c7_ab_leafshare_b5.log 2>&1
scripts/ab_variants.sh "leafshare6" --scenes=large --spp=16 > gpurun_out/c7_ab_leafshare_large.log 2>&1
tail -3 gpurun_out/c7_tests_leafshare.log
