#!/bin/bash
# round 2, GPU call 6: fast-shade tests + Whitted limit test, any-hit ordering A/B, ncu launch list of the final bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_shade_fast.py tests/test_gpu_integrators.py -m gpu -q -s) > gpurun_out/c6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c6_tests.log
scripts/ab_variants.sh "anyhit_unordered" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c6_ab_anyhit.log 2>&1
scripts/ab_variants.sh "anyhit_unordered" --scenes=large --spp=16 >> gpurun_out/c6_ab_anyhit.log 2>&1
python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_quick.json 2> gpurun_out/c6_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c6_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c6_ncu_launches.log 2>&1
tail -3 gpurun_out/c6_tests.log
