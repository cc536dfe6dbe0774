#!/bin/bash
# round 2, GPU call 12: sampler dimensions packed over the non-black lights -- same-path parity (GPU vs counter-driven restatement), smoke, timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_integrators.py tests/test_gpu_bvh_build.py -m gpu -q -s -k "not eight_seeds") > gpurun_out/c12_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c12_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c12_smoke.log
python scripts/time_scenes.py --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c12_time.log 2>&1
python bench.py --quick --config cornell --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench_cornell.json 2> gpurun_out/c12_bench_cornell.err
tail -3 gpurun_out/c12_tests.log; tail -2 gpurun_out/c12_smoke.log; cat gpurun_out/c12_time.log
