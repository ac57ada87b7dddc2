#!/bin/bash
# Round 2, GPU call 7: traversal stack top in a register + branch-free next-node selection (A/B against profiles/r02_sweep_float4_path_queue.log numbers)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal tests"; (time timeout 900 python -m pytest tests/test_gpu_traversal.py -q -m gpu) > $O/pytest_traversal.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_traversal.log | cut -c1-400
S=$O/sweep_r7.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "classroom 1280 720 64"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r7.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(40), 'ms %.2f  %.0f Msamples/s  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f launches %d'%(min(d['ms']),d['Msamples_s'],d['primary'],d['extend'],d['shade'],d['other'],d['tail'],d['launches']))
PY
echo "== one of 8 ranks"; timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1
