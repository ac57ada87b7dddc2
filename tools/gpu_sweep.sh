#!/bin/bash
# Scratch sweep: GPU tests + Dragon 1024x1024x256 timing for arena counts and register-cap / occupancy variants.
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
T="python tools/gpu_tune.py dragon 1024 1024 256 28"
for a in 1 2 4 8; do echo "arenas=$a"; B200PT_ARENAS=$a $T 2>&1 | tail -1; done
P=$PWD/monte-carlo-path-tracing_b200
for v in 3 5 6; do echo "min_ctas=$v"; B200PT_LIB=$P/build_v$v/libb200pt.so B200PT_CTAS_PER_SM=$v $T 2>&1 | tail -1; done
echo "1080p arenas=4"; python tools/gpu_tune.py dragon 1920 1080 512 30 2>&1 | tail -1
echo "1080p arenas=1"; B200PT_ARENAS=1 python tools/gpu_tune.py dragon 1920 1080 512 30 2>&1 | tail -1
echo "matpreview 128spp arenas=4"; python tools/gpu_tune.py matpreview 1024 1024 128 30 2>&1 | tail -1
echo "matpreview 128spp arenas=1"; B200PT_ARENAS=1 python tools/gpu_tune.py matpreview 1024 1024 128 30 2>&1 | tail -1
