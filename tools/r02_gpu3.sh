#!/bin/bash
# Round 2, GPU call 3: float4 radiance + CTA-aggregated k_bin_hits + spread tail paths; pointwise suites; create phases.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal + pointwise + host"; (time timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_pointwise.log | cut -c1-900
echo "== all other gpu tests"; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_traversal.py --deselect tests/test_gpu_pointwise.py --deselect tests/test_host_binary.py) > $O/pytest_gpu.log 2>&1; grep -E "^E  +Assertion|passed|failed|^FAILED" $O/pytest_gpu.log | cut -c1-400
S=$O/sweep_r3.log; : > $S
for sc in "dragon 1024 1024 256" "cornell-box 512 512 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "classroom 1280 720 64" "lte-orb-rough-glass 1024 1024 64"; do
  echo "## $sc" >> $S; timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S
done
cat $S
echo "== create phases"; B200PT_VERBOSE_CREATE=1 timeout 300 python tools/one_frame.py dragon 1024 1024 256 2>&1 | grep -E "b200pt create|render_ms" | tee $O/create_phases.log
echo "== one of 8 ranks"; for t in 32768 8192 131072; do B200PT_TAIL_PATHS=$t timeout 300 python tools/gpu_rank_breakdown.py 8 2>&1 | tail -1; done | tee $O/rank8_tail.log
echo "== bench b200"; timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'create', d['config']['scene_create_s'])
print({k:v for k,v in d['roofline'].items() if k not in ('kernels','formula','counters_source','bound_note')})
for k,v in d['configs'].items(): print(k, v.get('Msamples_s'), v.get('ms_per_step'), v.get('scene_create_s'), v.get('kernel_ms_single_arena'), v.get('error'))
PY
tail -3 $O/bench_n1.err
du -sh $O
