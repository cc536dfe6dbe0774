#!/bin/bash
# round 2, GPU call 3: new tests, the new bench.py, ncu launch list + full capture of the traversal kernels, node-loop unroll A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_full_size.py tests/test_gpu_guards.py tests/test_gpu_bvh_build.py tests/test_gpu_render.py -m gpu -x -q -s) > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c3_tests.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?" >> gpurun_out/c3_bench.err
scripts/ab_variants.sh "unroll2 unroll3" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c3_ab_unroll.log 2>&1
python scripts/profile_render.py bunny 8 > gpurun_out/c3_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_connect" -s 11 -c 4 -f -o gpurun_out/prof_bunny_r02 python scripts/profile_render.py bunny 8 > gpurun_out/c3_ncu_bunny.log 2>&1
python scripts/profile_render.py cornell 8 > gpurun_out/c3_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_connect|k_shade" -s 21 -c 6 -f -o gpurun_out/prof_cornell_r02 python scripts/profile_render.py cornell 8 > gpurun_out/c3_ncu_cornell.log 2>&1
python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_quick.json 2> gpurun_out/c3_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c3_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu_launches.log 2>&1
tail -3 gpurun_out/c3_tests.log; tail -2 gpurun_out/c3_bench.err
