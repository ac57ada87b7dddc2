#!/bin/bash
# Scratch sweep: GPU tests + timing of the BASELINE workloads (reduced spp for the long ones).
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "dragon"; python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
echo "matpreview 128spp"; B200PT_ARENAS=4 python tools/gpu_tune.py matpreview 1024 1024 128 30 2>&1 | tail -1
echo "volumetric 256spp"; B200PT_ARENAS=4 python tools/gpu_tune.py volumetric-caustic 1024 1024 256 30 2>&1 | tail -1
echo "cornell 1024spp"; python tools/gpu_tune.py cornell-box 512 512 1024 30 2>&1 | tail -1
