#!/bin/bash
# round 2, GPU call 16: early refill (node phase cut short when enough lanes have finished) -- sweep of the threshold
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for e in 0 8 12 16 20 24; do
  python scripts/time_scenes.py --scenes=bunny,cornell,glossy,large --spp=48 early_refill=$e
done > gpurun_out/c16_early_refill.log 2>&1
for e in "early_refill=12 refill_min=12" "early_refill=8 refill_min=8" "early_refill=16 refill_min=24" "early_refill=16 min_inner=12"; do
  python scripts/time_scenes.py --scenes=bunny,cornell --spp=48 $e
done >> gpurun_out/c16_early_refill.log 2>&1
cat gpurun_out/c16_early_refill.log
