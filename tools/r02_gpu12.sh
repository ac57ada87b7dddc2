#!/bin/bash
# Round 2, GPU call 12: pixel-granular visibility pre-pass (tests + timing, B200PT_PIXEL_CULL=0 = tiles only); 4 / 5 / 6 resident CTAs for the path shading variants.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
P=$PWD/monte-carlo-path-tracing_b200
echo "== parity tests"; (time timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu) > $O/pytest_parity.log 2>&1; grep -E "^E  +|passed|failed|^FAILED" $O/pytest_parity.log | cut -c1-500
S=$O/sweep_r12.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "cornell-box 512 512 256"; do
  echo "## $sc pixel-cull" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  echo "## $sc tiles-only" >> $S; B200PT_PIXEL_CULL=0 timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
for sc in "dragon 1024 1024 256" "matpreview 1024 1024 128" "lte-orb-rough-glass 1024 1024 64" "box 1024 1024 64" "cornell-box 512 512 256"; do
  for lib in libb200pt.so libb200pt_s5.so libb200pt_s6.so; do
    echo "## $sc $lib" >> $S; B200PT_LIB=$P/$lib timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r12.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(56), 'ms %.2f  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
echo "== one of 8 ranks"; timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1
