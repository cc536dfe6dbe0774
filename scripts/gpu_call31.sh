#!/bin/bash
# round 2, GPU call 31: the selector-based quantised walk on the tiny trees (Cornell, glossy) and the 5 M-triangle tree? (resident only up to 2^20 nodes) -- node_format 1 vs automatic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in -1 1 -1 1; do
  echo "== node_format $f"; python scripts/time_scenes.py --scenes=cornell,glossy --spp=48 node_format=$f
done > gpurun_out/c31_qn_tiny.log 2>&1
cat gpurun_out/c31_qn_tiny.log
