#!/bin/bash
# round 2, GPU call 9 (8 GPUs): the scaling sweep the driver runs -- N = 8, 4 (and 2) -- plus the single-process multi-GPU CLI on C5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c9_gpus.txt
for n in 8 4 2; do
  (time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3) > gpurun_out/c9_bench_n$n.json 2> gpurun_out/c9_bench_n$n.err
  echo "bench n=$n rc=$?" >> gpurun_out/c9_bench_n$n.err
done
(cd /tmp && time $OLDPWD/jet-pbrt_b200/jetpbrt 0 4096 3840 2160 0 0 --gpus 8) > gpurun_out/c9_cli_c5_gpus8.log 2>&1
python -m pytest tests/test_gpu_full_size.py::test_render_multi_on_two_gpus_equals_one -m gpu -q -s > gpurun_out/c9_test_multi.log 2>&1
tail -2 gpurun_out/c9_bench_n8.err; tail -4 gpurun_out/c9_cli_c5_gpus8.log
