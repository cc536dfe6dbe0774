#!/bin/bash
# round 2, GPU call 22: shading frames derived on the device, parallel host flatten -- parity suite, C3 upload / e2e timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x -k "not eight_seeds") > gpurun_out/c22_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c22_tests.log
JPBRT_FLATTEN_TIMING=1 python bench.py --quick --config large --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c22_bench_large.json 2> gpurun_out/c22_bench_large.err
tail -n 4 gpurun_out/c22_tests.log; grep flatten gpurun_out/c22_bench_large.err | head -8; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c22_bench_large.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","upload_s","bvh_build_s")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
PY
