#!/bin/bash
# round 2, GPU call 1: tests, peaks, traversal A/B (stack placement, node-loop unroll), ray-reordering sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
(time python -m pytest tests -m gpu -x -q -s) > gpurun_out/c1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c1_tests.log
jet-pbrt_b200/build/peaks_l2 > gpurun_out/c1_peaks.json 2> gpurun_out/c1_peaks.err
scripts/ab_variants.sh "r1 stk0 stk12 stk16 unroll2" --scenes=cornell,bunny,glossy,large --spp=16 > gpurun_out/c1_ab_stack.log 2>&1
for s in 0 3 4 5 6 19 20 21 22; do
  echo "== sort_rays=$s"
  python scripts/time_scenes.py sort_rays=$s --scenes=cornell,bunny,glossy,large --spp=16
done > gpurun_out/c1_ab_sort.log 2>&1
tail -3 gpurun_out/c1_tests.log
