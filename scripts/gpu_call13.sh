#!/bin/bash
# round 2, GPU call 13: traversal A/B -- lean child select + unguarded push, speculative (postponed-leaf) walk; parity of the speculative build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "lean spec speclean" --scenes=cornell,bunny,glossy,large --spp=48 > gpurun_out/c13_ab_trav.log 2>&1
cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so
cp jet-pbrt_b200/build/variants/speclean/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so
(time python -m pytest tests/test_gpu_render.py tests/test_gpu_parity.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c13_tests_speclean.log 2>&1
echo "tests rc=$?" >> gpurun_out/c13_tests_speclean.log
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
cat gpurun_out/c13_ab_trav.log; tail -5 gpurun_out/c13_tests_speclean.log
