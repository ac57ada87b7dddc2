#!/bin/bash
# Round 2, GPU call: warp packets (camera rays, first-vertex NEE rays) — tests and A/B timing; the parallel host SAH builder's
# phase times on the box's host cores.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal + pointwise + host"; (time timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_pointwise.log | cut -c1-900
S=$O/sweep_packets.log; : > $S
for sc in "dragon 1024 1024 256" "dragon 1920 1080 512" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "lte-orb-silver 1024 1024 64" "classroom 1280 720 64"; do
  for p in 0 1 3 7 15; do
    echo "## $sc B200PT_PACKETS=$p" >> $S; B200PT_PACKETS=$p timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
  done
done
cat $S
echo "== create phases"; B200PT_VERBOSE_CREATE=1 timeout 300 python tools/one_frame.py dragon 1024 1024 256 2>&1 | grep -E "b200pt create|render_ms" | tee $O/create_phases.log
echo "== one of 8 ranks"; for p in 0 3; do B200PT_PACKETS=$p timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; done | tee $O/rank8_packets.log
du -sh $O
