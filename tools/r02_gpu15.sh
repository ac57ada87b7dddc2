#!/bin/bash
# Round 2, GPU call 15: visibility pre-pass with the fine boxes grouped under the coarse ones and a parallel pixel list: tests, timing, kernel times.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== parity tests"; (time timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu) > $O/pytest_parity.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_parity.log | cut -c1-500
S=$O/sweep_r15.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "cornell-box 512 512 256" "volumetric-caustic 1024 1024 256" "matpreview 1024 1024 128"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r15.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(40), 'ms %.2f  %.0f Msamples/s  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['Msamples_s'],d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
for cfg in "dragon 1024 1024 256" "volumetric-caustic 1024 1024 64"; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(cull_tiles|scan_tiles|list_pixels)' --csv python tools/one_frame.py $cfg 2>/dev/null | grep -E "k_(cull|scan|list)" | awk -F'","' '{print $5, $NF}' | cut -c1-120
done
echo "== one of 8 ranks"; timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1
