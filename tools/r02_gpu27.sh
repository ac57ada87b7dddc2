#!/bin/bash
# Round 2, GPU call 27: scene-creation phases after the host-side changes, GPU tests, the bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/create_timing.py dragon > $O/create_phases.log 2>&1; grep -E "== create|BuildHostScene|UploadScene|geometry|SAH build|permutation|flatten" $O/create_phases.log | tail -8
echo "== pytest -m gpu"; (time timeout 1800 python -m pytest tests -m gpu -q) > $O/r02_pytest_gpu.log 2>&1; grep -E "passed|failed|^FAILED" $O/r02_pytest_gpu.log | tail -3
echo "== bench b200"; timeout 1500 python bench.py > $O/r02_bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], 'create', d['config']['scene_create_s'])
for k,v in d['configs'].items(): print(k, v.get('Msamples_s'), v.get('ms_per_step'), v.get('scene_create_s'), (v.get('parity') or {}).get('mean_ratio'), ((v.get('parity') or {}).get('exact_mode') or {}).get('pixels_within_2e-3'), v.get('error'))
PY
tail -2 $O/bench_n1.err
