#!/bin/bash
# Round 2, GPU call 22: exact-mode report after ordering the mesh-light CDFs like the reference's LBVH; replay tests; full GPU suite.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/replay_report.py --live 64 64 16 > $O/replay_report.log 2>&1; cut -c1-200 $O/replay_report.log
(time timeout 2400 python -m pytest tests -q -m gpu) > $O/pytest_gpu_r22.log 2>&1; grep -E "^E  +Assert|passed|failed|^FAILED" $O/pytest_gpu_r22.log | cut -c1-400
