#!/bin/bash
# round 2, GPU call 5: the whole GPU suite + smoke + the full default bench (unroll-2 kernels, fast-shade block)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q -s) > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c5_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c5_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c5_smoke.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
echo "bench rc=$?" >> gpurun_out/c5_bench.err
tail -4 gpurun_out/c5_tests.log; tail -2 gpurun_out/c5_smoke.log; tail -2 gpurun_out/c5_bench.err
