#!/bin/bash
# Round 2, GPU call 11: (a) one-triangle-per-pass leaf phase (lane thresholds 8 / 16) vs whole leaves; (b) 2 / 3 / 4 resident CTAs per SM for the
# ~120-register shading variants.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
P=$PWD/monte-carlo-path-tracing_b200
S=$O/sweep_r11.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "matpreview 1024 1024 128" "classroom 1280 720 64"; do
  for lib in libb200pt.so libb200pt_l8.so libb200pt_l16.so; do
    echo "## $sc $lib" >> $S; B200PT_LIB=$P/$lib timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
for sc in "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "lte-orb-rough-glass 1024 1024 64" "box 1024 1024 64" "material-testball 1280 720 64" "dining-room 1280 720 64"; do
  for lib in libb200pt.so libb200pt_v3.so libb200pt_v4.so; do
    echo "## $sc $lib" >> $S; B200PT_LIB=$P/$lib timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r11.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(56), 'ms %.2f  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
