#!/bin/bash
# Round 2, GPU call 18: end-of-launch straggler split in the persistent traversal: tests, timeline of one of 8 ranks, frame times.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== tests"; (time timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_parity.py -q -m gpu) > $O/pytest_split.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_split.log | cut -c1-500
B200PT_DUMP_TIMELINE=1 timeout 300 python tools/gpu_rank_breakdown.py 8 > $O/rank8_timeline_split.log 2>&1; grep "timeline" $O/rank8_timeline_split.log | grep -E "trace|tail" | awk '{print $3, $7}' | tr '\n' ';' | cut -c1-1500; echo; tail -1 $O/rank8_timeline_split.log
for w in 4 2; do timeout 300 python tools/gpu_rank_breakdown.py $w 2>&1 | tail -1; done
S=$O/sweep_r18.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "classroom 1280 720 64" "dining-room 1280 720 64"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_r18.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(40), 'ms %.2f  %.0f Msamples/s  prim %.2f ext %.2f shade %.2f other %.2f tail %.2f'%(min(d['ms']),d['Msamples_s'],d['primary'],d['extend'],d['shade'],d['other'],d['tail']))
PY
