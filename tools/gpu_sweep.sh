#!/bin/bash
cd "$(dirname "$0")/.."
P=$PWD/monte-carlo-path-tracing_b200
D="python tools/gpu_tune.py dragon 1024 1024 256 28"
echo "default"; $D 2>&1 | tail -1
echo "diffuse shade at 4 CTAs/SM (64 regs)"; B200PT_LIB=$P/build_s4/libb200pt.so $D 2>&1 | tail -1
echo "cornell default"; python tools/gpu_tune.py cornell-box 512 512 256 30 2>&1 | tail -1
echo "cornell 4 CTAs"; B200PT_LIB=$P/build_s4/libb200pt.so python tools/gpu_tune.py cornell-box 512 512 256 30 2>&1 | tail -1
