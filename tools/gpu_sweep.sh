#!/bin/bash
# Scratch sweep: GPU tests + Dragon 1024x1024x256 timing breakdown for a few launch tunables.
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "default"; python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
for v in 0 64; do echo "top=$v"; B200PT_TOP_NODES=$v python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1; done
echo "top=0 ctas=5"; B200PT_TOP_NODES=0 B200PT_CTAS_PER_SM=5 python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
