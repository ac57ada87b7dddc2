python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
for v in 4 12 16; do echo "min_inner=$v"; B200PT_MIN_INNER=$v python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1; done
for v in 1 512; do echo "top=$v"; B200PT_TOP_NODES=$v python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
