#!/bin/bash
# Round 2, GPU call 20: exact-mode (LCG replay) frames against the reference's exact fixtures.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/replay_report.py --live 64 64 16 > $O/replay_report.log 2>&1; cat $O/replay_report.log | cut -c1-400
