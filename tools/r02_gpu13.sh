#!/bin/bash
# Round 2, GPU call 13: (a) -fmad=false build (explicit fmaf() stays; no compiler contraction, ptxas cannot fuse mul+add differently per variant):
# bit-exactness tests + timing; (b) triangle vertices loaded without L1 allocation.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
P=$PWD/monte-carlo-path-tracing_b200
echo "== bit-exactness tests, -fmad=false library"; (B200PT_LIB=$P/libb200pt_nf.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_traversal.py tests/test_gpu_pointwise.py -q -m gpu) > $O/pytest_nf.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_nf.log | cut -c1-400
S=$O/sweep_r13.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "cornell-box 512 512 256" "classroom 1280 720 64"; do
  for lib in libb200pt.so libb200pt_nf.so libb200pt_na.so; do
    echo "## $sc $lib" >> $S; B200PT_LIB=$P/$lib timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r13.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(56), 'ms %.2f  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
