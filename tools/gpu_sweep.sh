#!/bin/bash
# Scratch sweep: GPU tests + BASELINE workloads + traversal tunables after the SEL / k_trace changes.
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
D="python tools/gpu_tune.py dragon 1024 1024 256 28"
echo "dragon"; $D 2>&1 | tail -1
echo "matpreview 128spp"; python tools/gpu_tune.py matpreview 1024 1024 128 30 2>&1 | tail -1
echo "volumetric 256spp"; python tools/gpu_tune.py volumetric-caustic 1024 1024 256 30 2>&1 | tail -1
for v in 4 12; do echo "min_inner=$v"; B200PT_MIN_INNER=$v $D 2>&1 | tail -1; done
for v in 8 24 32; do echo "refill=$v"; B200PT_REFILL=$v $D 2>&1 | tail -1; done
for v in 2 3 6 8; do echo "leaf=$v"; LEAF=$v $D 2>&1 | tail -1; done
