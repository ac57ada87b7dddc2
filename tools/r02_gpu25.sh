#!/bin/bash
# Round 2, GPU call 25: pointwise BSDF-sample mismatch shares per BSDF (200 000 probes each).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python tools/pointwise_report.py > $O/pointwise_report.log 2>&1; cut -c1-230 $O/pointwise_report.log | tail -120
