#!/bin/bash
# Round 2, GPU call 4: parity suites after the fixes; effect of warp-local live-entry compaction in k_shade and of k_settle.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal + pointwise + host"; (time timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_pointwise.log | cut -c1-500
echo "== all other gpu tests"; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_traversal.py --deselect tests/test_gpu_pointwise.py --deselect tests/test_host_binary.py) > $O/pytest_gpu.log 2>&1; grep -E "^E  +Assertion|passed|failed|^FAILED" $O/pytest_gpu.log | cut -c1-400
S=$O/sweep_shade_settle.log; : > $S
for sc in "dragon 1024 1024 256" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "dragon 1920 1080 512"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
echo "## one of 8 ranks (tools/gpu_rank_breakdown.py 8)" >> $S; timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -2 >> $S
cat $S
du -sh $O
