#!/bin/bash
cd "$(dirname "$0")/.."
P=$PWD/monte-carlo-path-tracing_b200
M="python tools/gpu_tune.py matpreview 1024 1024 64 30"
echo "matpreview default"; $M 2>&1 | tail -1
echo "matpreview conductor+generic at 3 CTAs/SM"; B200PT_LIB=$P/build_s3/libb200pt.so $M 2>&1 | tail -1
echo "dragon generic default"; B200PT_GENERIC_SHADE=1 python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
echo "dragon generic 3 CTAs"; B200PT_GENERIC_SHADE=1 B200PT_LIB=$P/build_s3/libb200pt.so python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
