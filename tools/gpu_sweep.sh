#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for t in 0 32768 131072 524288 2097152; do echo "tail=$t"; B200PT_TAIL_PATHS=$t python tools/gpu_rank_breakdown.py 1 2>&1 | tail -1; B200PT_TAIL_PATHS=$t python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; done
echo matpreview; for t in 0 131072; do B200PT_TAIL_PATHS=$t python tools/gpu_tune.py matpreview 1024 1024 64 30 2>&1 | tail -1; done
echo cornell; for t in 0 131072; do B200PT_TAIL_PATHS=$t python tools/gpu_tune.py cornell-box 512 512 256 30 2>&1 | tail -1; done
