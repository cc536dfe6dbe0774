#!/bin/bash
# round 2, GPU call 23: L1 / shared-memory carve-out preference of the wavefront kernels; 2-GPU sanity of the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cv in unset 0 25 50; do
  echo "== carveout $cv"
  if [ $cv = unset ]; then python scripts/time_scenes.py --scenes=bunny,cornell,large --spp=48; else JPBRT_CARVEOUT=$cv python scripts/time_scenes.py --scenes=bunny,cornell,large --spp=48; fi
done > gpurun_out/c23_carveout.log 2>&1
cat gpurun_out/c23_carveout.log
