#!/bin/bash
# round 2, GPU call 19: shade stage -- lights per trip of the NEE loop (ILP across lights), resident blocks per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "lu2 lu2b3 b3 b5" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c19_ab_shade_ilp.log 2>&1
cat gpurun_out/c19_ab_shade_ilp.log
