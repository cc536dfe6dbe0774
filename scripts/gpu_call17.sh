#!/bin/bash
# round 2, GPU call 17: light sample's normalised direction computed once (was three times per vertex and light) -- timing vs the previous build, parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "nocse" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c17_ab_light_cse.log 2>&1
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_integrators.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c17_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c17_tests.log
cat gpurun_out/c17_ab_light_cse.log; tail -5 gpurun_out/c17_tests.log
