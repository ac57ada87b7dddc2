#!/bin/bash
# Scratch sweep: GPU tests + speculative-descent variant of the traversal loop.
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
D="python tools/gpu_tune.py dragon 1024 1024 256 28"
P=$PWD/monte-carlo-path-tracing_b200
echo "dragon default"; $D 2>&1 | tail -1
echo "dragon speculative"; B200PT_LIB=$P/build_vs/libb200pt.so $D 2>&1 | tail -1
for v in 4 12 16; do echo "dragon speculative min_inner=$v"; B200PT_MIN_INNER=$v B200PT_LIB=$P/build_vs/libb200pt.so $D 2>&1 | tail -1; done
echo "matpreview default"; python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
echo "matpreview speculative"; B200PT_LIB=$P/build_vs/libb200pt.so python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
