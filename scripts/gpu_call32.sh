#!/bin/bash
# round 2, GPU call 32: final bench line + reference arm, ncu of the final kernels (bunny, bounces 0-1) and the launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err
echo "bench rc=$?" >> gpurun_out/c32_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/c32_bench_ref.json 2> gpurun_out/c32_bench_ref.err
python scripts/profile_render.py bunny 8 > gpurun_out/c32_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_connect|k_shade|k_logic" -s 37 -c 14 -f -o gpurun_out/prof_bunny_r02e python scripts/profile_render.py bunny 8 > gpurun_out/c32_ncu_bunny.log 2>&1
python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c32_bench_quick.json 2> gpurun_out/c32_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c32_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c32_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c32_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c32_smoke.log
tail -n 2 gpurun_out/c32_bench.err; tail -n 2 gpurun_out/c32_smoke.log
