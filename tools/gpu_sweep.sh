#!/bin/bash
cd "$(dirname "$0")/.."
for a in 4 8; do for w in 8 4 2; do B200PT_ARENAS=$a python tools/gpu_rank_breakdown.py $w 2>&1 | tail -1; done; done
for a in 2 3 4; do B200PT_ARENAS=$a python tools/gpu_rank_breakdown.py 1 2>&1 | tail -1; done
for v in 20 24 28; do echo "refill=$v"; B200PT_REFILL=$v python tools/gpu_rank_breakdown.py 1 2>&1 | tail -1;  B200PT_REFILL=$v python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; done
