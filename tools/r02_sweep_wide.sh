#!/bin/bash
# Round 2, first GPU call: correctness of the wide BVH path + first sweeps (layout, triangle postponing, packed fp32x2, TMA-staged top).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== traversal parity"; (time timeout 900 python -m pytest tests/test_gpu_traversal.py -x -q) > $O/pytest_traversal.log 2>&1; tail -5 $O/pytest_traversal.log
echo "== sweeps" 
S=$O/sweep_wide.log; : > $S
run() { echo "## $*" >> $S; env "$@" timeout 300 python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1 >> $S; }
run B200PT_BVH_LAYOUT=2
for t in 0 4 8 12 16 24; do run B200PT_BVH_LAYOUT=8 B200PT_TRI_MIN=$t; done
for r in 8 14 26; do run B200PT_BVH_LAYOUT=8 B200PT_REFILL=$r; done
X2=$PWD/monte-carlo-path-tracing_b200/libb200pt_x2.so
for t in 0 8 16; do run B200PT_LIB=$X2 B200PT_TRI_MIN=$t; done
for n in 73 256 585; do run B200PT_TOP_NODES=$n; done
run B200PT_CTAS_PER_SM=3
run B200PT_CTAS_PER_SM=2
for sc in "matpreview 1024 1024 128" "volumetric-caustic 1024 1024 256" "cornell-box 512 512 256"; do
  for l in 2 8; do echo "## $sc layout=$l" >> $S; B200PT_BVH_LAYOUT=$l timeout 300 python tools/gpu_tune.py $sc 28 2>&1 | tail -1 >> $S; done
done
cat $S
echo "== all gpu tests"; (time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
echo "== ncu --set full of the first launches of each kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(primary|trace|shade)' -c 5 -f -o /tmp/full_first5 \
    python tools/one_frame.py dragon 1024 1024 256 > $O/full_ncu.log 2>&1
python tools/ncu_summary.py /tmp/full_first5.ncu-rep > $O/full_first5_wide.txt 2>&1
ncu -i /tmp/full_first5.ncu-rep --page source --csv --kernel-name regex:k_trace --launch-count 1 2>/dev/null | cut -d, -f1-12 | gzip -9 > $O/k_trace_wide_source.csv.gz
du -sh $O
