#!/bin/bash
# Round 2, GPU call 23: exact mode at the full BASELINE sizes against the float16 reference frames.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python tools/replay_report.py --fullsize > $O/replay_fullsize.log 2>&1; cut -c1-250 $O/replay_fullsize.log
