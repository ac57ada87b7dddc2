#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "dragon sah"; python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
echo "dragon lbvh (GPU flatten)"; B200PT_BVH_BUILDER=lbvh python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
echo "matpreview lbvh"; B200PT_BVH_BUILDER=lbvh python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
