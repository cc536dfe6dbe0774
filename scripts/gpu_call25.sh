#!/bin/bash
# round 2, GPU call 25: 32-byte quantised BVH nodes (one 256-bit load per node step instead of two) -- timing and parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "qn" --scenes=bunny,cornell,glossy,large --spp=48 > gpurun_out/c25_ab_qnodes.log 2>&1
for v in base qn; do
  if [ $v = qn ]; then cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so; cp jet-pbrt_b200/build/variants/qn/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so; fi
  echo "== $v counts"; python scripts/time_scenes.py --scenes=bunny,large --spp=8 count_traversal=1 2>&1 | tail -n 2
done >> gpurun_out/c25_ab_qnodes.log 2>&1
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_guards.py tests/test_gpu_bvh_build.py tests/test_gpu_integrators.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c25_tests_qn.log 2>&1
echo "tests rc=$?" >> gpurun_out/c25_tests_qn.log
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
cat gpurun_out/c25_ab_qnodes.log; tail -n 5 gpurun_out/c25_tests_qn.log
