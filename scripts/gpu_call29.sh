#!/bin/bash
# round 2, GPU call 29: whole GPU suite, smoke, full bench + reference arm, then ncu (launch list of the bench command; full capture of traversal + shade kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -s) > gpurun_out/c29_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c29_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c29_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c29_smoke.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
echo "bench rc=$?" >> gpurun_out/c29_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/c29_bench_ref.json 2> gpurun_out/c29_bench_ref.err
python scripts/profile_render.py bunny 8 > gpurun_out/c29_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_connect|k_shade|k_logic" -s 37 -c 14 -f -o gpurun_out/prof_bunny_r02d python scripts/profile_render.py bunny 8 > gpurun_out/c29_ncu_bunny.log 2>&1
python scripts/profile_render.py cornell 8 > gpurun_out/c29_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_connect|k_shade" -s 21 -c 6 -f -o gpurun_out/prof_cornell_r02d python scripts/profile_render.py cornell 8 > gpurun_out/c29_ncu_cornell.log 2>&1
python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c29_bench_quick.json 2> gpurun_out/c29_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c29_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c29_ncu_launches.log 2>&1
tail -n 3 gpurun_out/c29_tests.log; tail -n 2 gpurun_out/c29_bench.err; cat gpurun_out/c29_smoke.log | tail -n 2
for o in "refill_min=12" "refill_min=20" "min_inner=6" "min_inner=10" "min_inner=12 refill_min=20"; do python scripts/time_scenes.py --scenes=bunny --spp=48 $o; done > gpurun_out/c29_tune_qn.log 2>&1
cat gpurun_out/c29_tune_qn.log
