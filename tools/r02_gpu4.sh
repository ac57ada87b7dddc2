#!/bin/bash
# Round 2, GPU call 4: pointwise suites after the test fixes; L2 persistence window over the tree and push-prefetch variants.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal + pointwise + host"; (time timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_pointwise.log | cut -c1-900
S=$O/sweep_l2_prefetch.log; : > $S
run() { echo "## $*" >> $S; env "${@:5}" timeout 300 python tools/gpu_tune.py $1 $2 $3 $4 28 2>&1 | tail -1 >> $S; }
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "matpreview 1024 1024 128" "classroom 1280 720 64"; do
  run $sc B200PT_L2_PERSIST_MB=0
  run $sc B200PT_L2_PERSIST_MB=32
  run $sc B200PT_L2_PERSIST_MB=72
  run $sc B200PT_L2_PERSIST_MB=100
  run $sc B200PT_PUSH_PREFETCH=1
  run $sc B200PT_PUSH_PREFETCH=1 B200PT_L2_PERSIST_MB=72
done
python - <<'PY'
import json
cur=None
for l in open('gpurun_out/sweep_l2_prefetch.log'):
    if l.startswith('## '): cur=l[3:].strip()
    elif l.startswith('{"cap'):
        d=json.loads(l); print(cur.ljust(70), 'ms %.2f prim %.2f ext %.2f shade %.2f other %.2f'%(min(d['ms']),d['primary'],d['extend'],d['shade'],d['other']))
PY
du -sh $O
