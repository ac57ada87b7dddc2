#!/bin/bash
# Round 2, GPU call: parity suites (incl. the six extra reference scenes), wide-tree diagnostic, bench line with the C1/C3/C4/C5
# sub-records, per-kernel ncu counters of C2 / C3 / C4 (reduced spp for C3 / C4: the counters are rates).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== traversal + pointwise + host"; (time timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; grep -E "^E  +Assertion|passed|failed" $O/pytest_pointwise.log | cut -c1-900
echo "== diag"; timeout 300 python tools/r02_diag_wide.py > $O/diag_wide.log 2>&1; tail -40 $O/diag_wide.log | cut -c1-600
echo "== all other gpu tests"; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_traversal.py --deselect tests/test_gpu_pointwise.py --deselect tests/test_host_binary.py) > $O/pytest_gpu.log 2>&1; grep -E "^E  +Assertion|passed|failed|^FAILED" $O/pytest_gpu.log | cut -c1-400
echo "== bench b200"; timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 3000 $O/bench_n1.json; tail -5 $O/bench_n1.err
M=$(python tools/ncu_counters.py --metrics)
for cfg in "dragon 1024 1024 256" "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256"; do
  set -- $cfg
  echo "== ncu counters $cfg"
  timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/raw_$1.csv python tools/one_frame.py $cfg > $O/one_frame_$1.log 2>&1
  python tools/ncu_counters.py $O/raw_$1.csv $O/r02_counters_$1_$2x$3x$4.json "$1 $2x$3x$4" 2>&1 | head -14
  gzip -9f $O/raw_$1.csv
done
du -sh $O
