#!/bin/bash
# round 2, GPU call 10: relaxed-shade tests, BSDF parity (Lambda(wo) hoist must stay bit-exact), final bench, shade_math A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_shade_fast.py tests/test_gpu_parity.py tests/test_gpu_integrators.py tests/test_gpu_render.py -m gpu -q -s) > gpurun_out/c10_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c10_tests.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
echo "bench rc=$?" >> gpurun_out/c10_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/c10_bench_ref.json 2> gpurun_out/c10_bench_ref.err
for o in "shade_math=0" "shade_math=1"; do
  echo "== $o"
  python scripts/time_scenes.py $o --scenes=cornell,bunny,glossy --spp=48
done > gpurun_out/c10_ab_shade_math.log 2>&1
tail -3 gpurun_out/c10_tests.log; tail -2 gpurun_out/c10_bench.err
