python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1
for ct in 0.25 4; do echo "ct=$ct"; B200PT_SAH_TRAVERSAL_COST=$ct python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
