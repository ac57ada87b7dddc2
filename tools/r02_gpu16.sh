#!/bin/bash
# Round 2, GPU call 16: compute-sanitizer (memcheck, racecheck, initcheck) over small frames of scenes that together reach every kernel.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck; do
  for cfg in "dragon 96 96 8" "matpreview 64 64 8" "volumetric-caustic 64 64 8" "cornell-box 64 64 8"; do
    echo "== $tool $cfg"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/one_frame.py $cfg > $O/sanitizer_${tool}_${cfg%% *}.log 2>&1
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|render_ms" $O/sanitizer_${tool}_${cfg%% *}.log | tail -2
  done
done
