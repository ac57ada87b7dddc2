#!/bin/bash
# Round 2, 4-GPU call: bench.py as the driver launches it at N=4 (C2 tile split + the C5 sub-record).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','steps')}, 'e2e', d['e2e']['value'])
for k,v in d['configs'].items(): print(k, v.get('n_gpus'), v.get('Msamples_s'), v.get('ms_per_step'), v.get('kernel_ms_single_arena'), (v.get('parity') or {}).get('mean_ratio'), v.get('error'))
PY
tail -3 $O/bench_n4.err
