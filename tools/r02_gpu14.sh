#!/bin/bash
# Round 2, GPU call 14: contraction only in the front end (-Xptxas -fmad=false): do the shading variants, compiled under different register
# caps, agree bit for bit again?  Timing against the default and against -fmad=false.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
P=$PWD/monte-carlo-path-tracing_b200
echo "== bit-exactness tests, -Xptxas -fmad=false library"; (B200PT_LIB=$P/libb200pt_pf.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_traversal.py -q -m gpu) > $O/pytest_pf.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_pf.log | cut -c1-400
S=$O/sweep_r14.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "cornell-box 512 512 256"; do
  for lib in libb200pt.so libb200pt_pf.so libb200pt_nf.so; do
    echo "## $sc $lib" >> $S; B200PT_LIB=$P/$lib timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r14.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(56), 'ms %.2f  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
