#!/bin/bash
# round 2, GPU call 14: shadow-record fetch variants of k_connect (dependent loads / parallel / L2 prefetch / cp.async staging), shade prefetch; full GPU suite on the new base
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ab_variants.sh "io0 io2 io3 shpf" --scenes=cornell,bunny,glossy --spp=48 > gpurun_out/c14_ab_io.log 2>&1
(time python -m pytest tests -m gpu -q -x -k "not eight_seeds") > gpurun_out/c14_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c14_tests.log
cat gpurun_out/c14_ab_io.log; tail -5 gpurun_out/c14_tests.log
