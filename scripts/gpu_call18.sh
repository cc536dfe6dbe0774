#!/bin/bash
# round 2, GPU call 18: ray indices claimed one 32-ray chunk ahead (the work cursor's atomic off the critical path) vs exact claims per refill
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "noclaim" --scenes=cornell,bunny,glossy,large --spp=48 > gpurun_out/c18_ab_claim.log 2>&1
python scripts/time_scenes.py --scenes=cornell,bunny,glossy --spp=48 trav_blocks=5 >> gpurun_out/c18_ab_claim.log 2>&1
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_integrators.py tests/test_gpu_guards.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c18_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c18_tests.log
cat gpurun_out/c18_ab_claim.log; tail -5 gpurun_out/c18_tests.log
