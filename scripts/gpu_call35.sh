#!/bin/bash
# round 2, GPU call 35: node steps per warp vote re-measured now that the quantised walk is issue-bound; min_inner / refill_min
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "un3 un4" --scenes=bunny,cornell --spp=48 > gpurun_out/c35_ab_unroll.log 2>&1
for o in "min_inner=10" "min_inner=10 refill_min=12" "min_inner=12"; do python scripts/time_scenes.py --scenes=bunny --spp=48 $o; done >> gpurun_out/c35_ab_unroll.log 2>&1
cat gpurun_out/c35_ab_unroll.log
