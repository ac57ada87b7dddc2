#!/bin/bash
# Round-2 final evidence run (one gpurun call): GPU tests, smoke, both bench arms, ncu launch list of the bench command.  The ncu
# counters / --set full captures of tools/gpu_evidence_r02.sh still describe the kernels (nothing device-side changed since).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu"; (time timeout 1800 python -m pytest tests -m gpu -q) > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/bench_reference.err; cut -c1-400 $O/r02_bench_reference.json
echo "== bench b200"; timeout 1500 python bench.py > $O/r02_bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], 'create', d['config']['scene_create_s'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('kernels','formula','counters_source','bound_note')})
print('cpu', d['cpu_baseline']); print('parity', d['parity'])
for k,v in d['configs'].items(): print(k, v.get('Msamples_s'), v.get('ms_per_step'), v.get('scene_create_s'), (v.get('parity') or {}).get('mean_ratio'), (v.get('parity') or {}).get('exact_mode'), v.get('error'))
PY
tail -3 $O/bench_n1.err
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/bench_under_ncu.log 2>&1
grep -c k_trace $O/r02_launches_bench.csv
gzip -9f $O/r02_launches_bench.csv
du -sh $O
