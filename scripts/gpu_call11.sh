#!/bin/bash
# round 2, GPU call 11: tree-quality knobs on the final kernels (reinsertion passes, SAH traversal cost, leaf size) + full suite after the shade-mode removal
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 2 4 8; do
  echo "== reinsert=$p"
  JPBRT_BVH_REINSERT=$p python scripts/time_scenes.py count_traversal=0 --scenes=bunny --spp=48
done > gpurun_out/c11_ab_reinsert.log 2>&1
for t in 50 100 150 200; do for l in 2 4 6; do
  echo "== trav=$t leaf=$l"
  JPBRT_BVH_TRAV=$t JPBRT_BVH_LEAF=$l python scripts/time_scenes.py --scenes=bunny,cornell --spp=48
done; done > gpurun_out/c11_ab_sah.log 2>&1
(time python -m pytest tests -m gpu -q) > gpurun_out/c11_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c11_tests.log
tail -3 gpurun_out/c11_tests.log
