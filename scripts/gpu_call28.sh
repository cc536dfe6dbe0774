#!/bin/bash
# round 2, GPU call 28: triangle normal in the slot's spare quarter (new base), newest stack entry cached in a register (variant tos) -- timing, parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "tos" --scenes=bunny,cornell,glossy,large --spp=48 > gpurun_out/c28_ab_tos.log 2>&1
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_bvh_build.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c28_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c28_tests.log
cat gpurun_out/c28_ab_tos.log; tail -n 5 gpurun_out/c28_tests.log
