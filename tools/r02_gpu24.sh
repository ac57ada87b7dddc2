#!/bin/bash
# Round 2, GPU call 24: classify the first divergence of the worst exact-mode pixels (state / hit distance / radiance / throughput).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/replay_trace2.log; : > $L
for sc in "box 24 24 4" "dining-room 24 24 4" "classroom 24 24 4" "synthetic_dielectrics_conductor_cylinder 32 32 4" "synthetic_bump_bitmap_mesh_disk 32 32 4" "lte-orb-rough-glass 24 24 4" "cornell-box 24 24 4" "dragon 64 64 16"; do
  timeout 600 python tools/replay_trace.py $sc 5 >> $L 2>&1
done
cut -c1-260 $L
