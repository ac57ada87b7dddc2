#!/bin/bash
# Round 2, GPU call 3: pointwise parity suites, whole GPU suite, binary-vs-wide instruction counts under ncu (same launch, one arena).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== traversal + pointwise"; (time timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_pointwise.py tests/test_host_binary.py -q -m gpu) > $O/pytest_pointwise.log 2>&1; tail -40 $O/pytest_pointwise.log | cut -c1-400
echo "== all gpu tests"; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_traversal.py --deselect tests/test_gpu_pointwise.py --deselect tests/test_host_binary.py) > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log | cut -c1-300
S=$O/sweep_layout.log; : > $S
for l in 2 8; do echo "## layout=$l" >> $S; B200PT_BVH_LAYOUT=$l timeout 300 python tools/gpu_tune.py dragon 1024 1024 256 28 2>&1 | tail -1 >> $S; done
cat $S
for l in 2 8; do
  B200PT_ARENAS=1 B200PT_BVH_LAYOUT=$l timeout 600 ncu --set full --clock-control none -k regex:'k_trace' -c 2 -f -o /tmp/trace_l$l python tools/one_frame.py dragon 1024 1024 256 > $O/ncu_l$l.log 2>&1
  python tools/ncu_summary.py /tmp/trace_l$l.ncu-rep > $O/ncu_k_trace_layout$l.txt 2>&1
done
grep -E "^---|inst_executed.sum|thread_inst_executed_per|gpu__time_duration|issue_active" $O/ncu_k_trace_layout*.txt
du -sh $O
