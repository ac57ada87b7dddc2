#!/bin/bash
# Round-2 evidence run (one gpurun call): GPU tests, smoke, both bench arms, ncu launch list of the bench command, per-kernel counters of
# C2 / C3 / C4, one --set full capture with SASS source pages.  Only text / compressed CSV travels back (gpurun_out is capped at 64 MiB).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu"; (time timeout 1800 python -m pytest tests -m gpu -q) > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/bench_reference.err; cut -c1-600 $O/r02_bench_reference.json
echo "== bench b200"; timeout 1200 python bench.py > $O/r02_bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], 'create', d['config']['scene_create_s'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('kernels','formula','counters_source','bound_note')})
print('cpu', d['cpu_baseline']); print('refgpu', d['reference_gpu_baseline'])
for k,v in d['configs'].items(): print(k, v.get('Msamples_s'), v.get('ms_per_step'), v.get('scene_create_s'), v.get('kernel_ms_single_arena'), (v.get('parity') or {}).get('mean_ratio'), v.get('error'))
PY
tail -3 $O/bench_n1.err
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/bench_under_ncu.log 2>&1
grep -c k_trace $O/r02_launches_bench.csv
M=$(python tools/ncu_counters.py --metrics)
for cfg in "dragon 1024 1024 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256"; do
  set -- $cfg
  echo "== ncu counters $cfg"
  timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/raw_$1.csv python tools/one_frame.py $cfg > $O/one_frame_$1.log 2>&1
  python tools/ncu_counters.py $O/raw_$1.csv $O/r02_counters_$1_$2x$3x$4.json "$1 $2x$3x$4" 2>&1 | head -9
  rm -f $O/raw_$1.csv
done
echo "== ncu --set full of the first launches"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(primary|trace|shade)' -c 4 -f -o /tmp/full_r02 \
    python tools/one_frame.py dragon 1024 1024 256 > $O/full_ncu.log 2>&1
python tools/ncu_summary.py /tmp/full_r02.ncu-rep > $O/r02_full_first4.txt 2>&1
python tools/ncu_source.py /tmp/full_r02.ncu-rep regex:k_trace $O/r02_k_trace_source.csv.gz 0 | cut -c1-80
python tools/ncu_source.py /tmp/full_r02.ncu-rep regex:k_primary $O/r02_k_primary_source.csv.gz 0 | cut -c1-80
python tools/ncu_source.py /tmp/full_r02.ncu-rep regex:k_shade $O/r02_k_shade_source.csv.gz 0 | cut -c1-80
gzip -9f $O/r02_launches_bench.csv
du -sh $O
