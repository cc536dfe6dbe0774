#!/bin/bash
# On the GPU box: time the BASELINE scenes with the in-tree library and with each A/B variant built by scripts/build_variant.sh.
# usage: scripts/ab_variants.sh "variant1 variant2 ..." [time_scenes args]
cd "$(dirname "$0")/.."
variants=$1; shift
cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so
for v in base $variants; do
  if [ $v = base ]; then cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so; else cp jet-pbrt_b200/build/variants/$v/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so; fi
  echo "== $v"
  python scripts/time_scenes.py "$@"
done
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
