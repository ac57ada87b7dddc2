#!/bin/bash
cd "$(dirname "$0")/.."
for b in 0 1; do echo "bin_single=$b"; B200PT_BIN_SINGLE=$b python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1; B200PT_BIN_SINGLE=$b python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; B200PT_BIN_SINGLE=$b python tools/gpu_tune.py cornell-box 512 512 256 30 2>&1 | tail -1; done
