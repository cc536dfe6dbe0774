#!/bin/bash
# round 2, GPU call 26: closest hits / occlusion of the quantised-node build against the fp32-node build on 4 x 2^20 queries per scene
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in bunny large; do python scripts/dump_queries.py $s /tmp/base_$s.npz; done > gpurun_out/c26_queries.log 2>&1
cp jet-pbrt_b200/libjetpbrt_b200.so /tmp/base.so; cp jet-pbrt_b200/build/variants/qn/libjetpbrt_b200.so jet-pbrt_b200/libjetpbrt_b200.so
for s in bunny large; do python scripts/dump_queries.py $s /tmp/qn_$s.npz; done >> gpurun_out/c26_queries.log 2>&1
cp /tmp/base.so jet-pbrt_b200/libjetpbrt_b200.so
python - >> gpurun_out/c26_queries.log 2>&1 <<'PY'
import numpy as np
for s in ("bunny", "large"):
    a, b = np.load(f"/tmp/base_{s}.npz"), np.load(f"/tmp/qn_{s}.npz")
    for k in a.files:
        d = a[k] != b[k]
        print(s, k, "differ", int(d.sum()), "of", d.size)
        if d.sum() and k.startswith("prim"):
            i = np.where(d)[0][:6]
            print("   base", a[k][i], "qn", b[k][i], "t base", a["t" + k[4:]][i], "t qn", b["t" + k[4:]][i])
PY
cat gpurun_out/c26_queries.log
