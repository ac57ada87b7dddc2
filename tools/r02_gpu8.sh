#!/bin/bash
# Round 2, GPU call 8: re-tune of the persistent loop's two thresholds after this round's changes (refill group size, inner-phase exit).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
S=$O/sweep_r8.log; : > $S
for sc in "dragon 1024 1024 256" "matpreview 1024 1024 128"; do
  for r in 12 16 20 24 28; do for m in 4 8 12; do
    echo "## $sc B200PT_REFILL=$r B200PT_MIN_INNER=$m" >> $S; B200PT_REFILL=$r B200PT_MIN_INNER=$m timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done; done
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r8.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(64), 'ms %.2f  prim %.2f ext %.2f shade %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade']))
PY
