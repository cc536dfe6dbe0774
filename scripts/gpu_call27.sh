#!/bin/bash
# round 2, GPU call 27: node format as a per-scene choice (option node_format) -- parity suite with the automatic choice, A/B by option
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x -k "not eight_seeds") > gpurun_out/c27_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c27_tests.log
for f in 0 -1 1; do
  echo "== node_format $f"; python scripts/time_scenes.py --scenes=bunny,cornell,glossy,large --spp=48 node_format=$f
done > gpurun_out/c27_ab_node_format.log 2>&1
cat gpurun_out/c27_ab_node_format.log; tail -n 5 gpurun_out/c27_tests.log
