#!/bin/bash
cd "$(dirname "$0")/.."
for t in 0 16384 131072 1048576 1073741824; do echo "prefetch_below=$t"; B200PT_PREFETCH_BELOW=$t python tools/gpu_rank_breakdown.py 1 2>&1 | tail -1; B200PT_PREFETCH_BELOW=$t python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; done
