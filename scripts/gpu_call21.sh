#!/bin/bash
# round 2, GPU call 21: k_logic -- prefetch of the whole fetch's records, resident blocks per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "nolpf lb3 lb6" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c21_ab_logic.log 2>&1
cat gpurun_out/c21_ab_logic.log
