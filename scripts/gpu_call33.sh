#!/bin/bash
# round 2, GPU call 33 (2 GPUs): multi-GPU paths with the final build -- jpbrt_render_multi test, torchrun bench at N = 2 (weak + strong + reduce_check)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_full_size.py -m gpu -q -k "multi") > gpurun_out/c33_test_multi.log 2>&1
echo "tests rc=$?" >> gpurun_out/c33_test_multi.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/c33_bench_n2.json 2> gpurun_out/c33_bench_n2.err
echo "bench rc=$?" >> gpurun_out/c33_bench_n2.err
tail -n 4 gpurun_out/c33_test_multi.log; tail -n 3 gpurun_out/c33_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c33_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","scaling")}, "e2e", d["e2e"]["value"], "reduce_check", d.get("reduce_check"))
for k,v in d.get("strong",{}).items(): print(k, {kk:v.get(kk) for kk in ("value","ms_per_step","spp_total")}, v.get("reduce_alone_ms"))
PY
