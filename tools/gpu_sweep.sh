#!/bin/bash
cd "$(dirname "$0")/.."
P=$PWD/monte-carlo-path-tracing_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "variant tests"; B200PT_LIB=$P/build_sp/libb200pt.so python -m pytest tests -m gpu -x -q 2>&1 | tail -4
D="python tools/gpu_tune.py dragon 1024 1024 256 28"
echo "default"; $D 2>&1 | tail -1
for v in 8 12 16 20 24; do echo "spare refill=$v"; B200PT_REFILL=$v B200PT_LIB=$P/build_sp/libb200pt.so $D 2>&1 | tail -1; done
echo "spare refill=16 min_inner=12"; B200PT_MIN_INNER=12 B200PT_REFILL=16 B200PT_LIB=$P/build_sp/libb200pt.so $D 2>&1 | tail -1
echo "matpreview default"; python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
echo "matpreview spare"; B200PT_REFILL=16 B200PT_LIB=$P/build_sp/libb200pt.so python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1
