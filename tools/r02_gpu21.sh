#!/bin/bash
# Round 2, GPU call 21: exact-mode traces of the worst pixels on the scenes that do not agree per pixel yet.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/replay_trace.log; : > $L
for sc in "synthetic_dielectrics_conductor_cylinder 32 32 4" "synthetic_isotropic_medium_null_surface 32 32 4" "volumetric-caustic 24 24 4" "synthetic_bump_bitmap_mesh_disk 32 32 4" "box 24 24 4" "classroom 24 24 4"; do
  timeout 600 python tools/replay_trace.py $sc 2 >> $L 2>&1
done
cut -c1-330 $L
