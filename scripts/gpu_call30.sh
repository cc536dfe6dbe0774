#!/bin/bash
# round 2, GPU call 30: quantised walk with per-ray PRMT selectors for the near / far planes (new base) vs both distances + min / max (qnold); parity; bench line with the 32-byte gather ceiling
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "qnold" --scenes=bunny,cornell --spp=48 > gpurun_out/c30_ab_qnselect.log 2>&1
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_guards.py tests/test_gpu_bvh_build.py tests/test_gpu_integrators.py tests/test_abi.py -m gpu -q -x -k "not eight_seeds") > gpurun_out/c30_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c30_tests.log
python bench.py --quick --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c30_bench_quick.json 2> gpurun_out/c30_bench_quick.err
cat gpurun_out/c30_ab_qnselect.log; tail -n 5 gpurun_out/c30_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c30_bench_quick.json').read().strip().splitlines()[-1])
r=d["roofline"]; print(d["value"], {k:r.get(k) for k in ("achieved","peak","frac","node_bytes","distinct_bytes_per_ray","gather_peak_32_byte_records_gbs","gather_peak_64_byte_records_gbs")})
PY
