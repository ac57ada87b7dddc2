#!/bin/bash
# Scratch sweep: GPU tests + GPU LBVH builder vs host SAH builder (build time, render time).
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
D="python tools/gpu_tune.py dragon 1024 1024 256 28"
echo "dragon sah"; $D 2>&1 | tail -1
for v in 2 4 8; do echo "dragon lbvh leaf=$v"; B200PT_BVH_BUILDER=lbvh LEAF=$v $D 2>&1 | tail -1; done
echo "matpreview sah"; python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
echo "matpreview lbvh"; B200PT_BVH_BUILDER=lbvh python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
