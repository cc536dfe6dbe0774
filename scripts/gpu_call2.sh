#!/bin/bash
# round 2, GPU call 2: traversal A/B (round-1 kernels vs shared-memory stack depth 0/8/12/16, node-loop unroll), tunables of the new kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "r1 stk0 stk12 stk16 unroll2" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c2_ab_stack.log 2>&1
scripts/ab_variants.sh "r1" --scenes=large --spp=16 > gpurun_out/c2_ab_stack_large.log 2>&1
for o in "trav_blocks=5" "min_inner=4" "min_inner=12" "refill_min=8" "refill_min=24" "min_inner=6 refill_min=12"; do
  echo "== $o"
  python scripts/time_scenes.py $o --scenes=cornell,bunny,glossy --spp=48
done > gpurun_out/c2_ab_tune.log 2>&1
tail -2 gpurun_out/c2_ab_tune.log
