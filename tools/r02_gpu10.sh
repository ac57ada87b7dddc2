#!/bin/bash
# Round 2, GPU call 10: the ~120-register shading variants compiled for 3 resident CTAs per SM (80 registers) vs left to the compiler (2 CTAs).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
S=$O/sweep_r10.log; : > $S
for sc in "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "lte-orb-rough-glass 1024 1024 64" "classroom 1280 720 64" "box 1024 1024 64" "material-testball 1280 720 64"; do
  echo "## $sc default" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  echo "## $sc 3ctas" >> $S; B200PT_LIB=$PWD/monte-carlo-path-tracing_b200/libb200pt_v3.so timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r10.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(48), 'ms %.2f  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
echo "== tail specialisation: one of 8 ranks"; timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1
