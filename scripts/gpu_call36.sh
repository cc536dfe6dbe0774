#!/bin/bash
# round 2, GPU call 36 (last): whole GPU suite, smoke, final bench line with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/c36_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c36_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c36_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c36_smoke.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err
echo "bench rc=$?" >> gpurun_out/c36_bench.err
tail -n 6 gpurun_out/c36_tests.log; cat gpurun_out/c36_smoke.log; tail -n 2 gpurun_out/c36_bench.err
