#!/bin/bash
# Round 2, GPU call 26: tail kernel with two lanes per path (the helper lane traces the NEE ray while the owner extends the path).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== tests"; (time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_traversal.py -q -m gpu -x) > $O/pytest_tail2.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_tail2.log | cut -c1-400
B200PT_DUMP_TIMELINE=1 timeout 300 python tools/gpu_rank_breakdown.py 8 > $O/rank8_timeline_tail2.log 2>&1; grep "timeline" $O/rank8_timeline_tail2.log | grep -E "tail" | awk '{print $7}' | sort -n | tail -2 | tr '\n' ' '; echo; tail -1 $O/rank8_timeline_tail2.log
for w in 4 2; do timeout 300 python tools/gpu_rank_breakdown.py $w 2>&1 | tail -1; done
S=$O/sweep_r26.log; : > $S
for sc in "dragon 1024 1024 256" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "classroom 1280 720 64"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r26.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(40), 'ms %.2f  %.0f Msamples/s  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['Msamples_s'],d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
