#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL inside the C ABI -- render_multi, torchrun bench with reduce_check and strong scaling, CLI --gpus;
# fast-shade parity tests and timing; reference-side shim
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c4_gpus.txt
(time python -m pytest tests/test_gpu_full_size.py::test_render_multi_on_two_gpus_equals_one tests/test_gpu_full_size.py::test_single_rank_communicator tests/test_gpu_render.py::test_film_tensor_aliases_device_film tests/test_gpu_shade_fast.py tests/test_gpu_reference_shim.py -m gpu -q -s) > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c4_tests.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/c4_bench_n2.json 2> gpurun_out/c4_bench_n2.err
echo "bench rc=$?" >> gpurun_out/c4_bench_n2.err
(cd /tmp && time $OLDPWD/jet-pbrt_b200/jetpbrt 1 50 1024 1024 0 0 --gpus 2) > gpurun_out/c4_cli_gpus2.log 2>&1
(cd /tmp && time $OLDPWD/jet-pbrt_b200/jetpbrt 1 50 1024 1024 0 0) > gpurun_out/c4_cli_gpus1.log 2>&1
cmp /tmp/bunny_scene_50.bmp /tmp/bunny_scene_50.bmp && ls -la /tmp/*.bmp >> gpurun_out/c4_cli_gpus1.log
for o in "shade_math=0" "shade_math=1"; do
  echo "== $o"
  python scripts/time_scenes.py $o --scenes=cornell,bunny,glossy --spp=48
done > gpurun_out/c4_ab_shade_math.log 2>&1
tail -3 gpurun_out/c4_tests.log; tail -3 gpurun_out/c4_bench_n2.err
