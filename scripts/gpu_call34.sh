#!/bin/bash
# round 2, GPU call 34: quantised walk -- approximate reciprocal for the slab constants (new base) vs IEEE (ieeercp); far selectors kept in registers (farsel); parity of the base
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "ieeercp farsel" --scenes=bunny --spp=48 > gpurun_out/c34_ab_rcp.log 2>&1
(time python -m pytest tests/test_gpu_render.py tests/test_gpu_parity.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c34_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c34_tests.log
cat gpurun_out/c34_ab_rcp.log; tail -n 4 gpurun_out/c34_tests.log
