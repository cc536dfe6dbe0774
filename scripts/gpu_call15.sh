#!/bin/bash
# round 2, GPU call 15: guided self-scheduling of the shade-stage work fetches (fetch size shrinks as the queue drains)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "g0 g32 g128" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c15_ab_guided.log 2>&1
cat gpurun_out/c15_ab_guided.log
