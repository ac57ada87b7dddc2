#!/bin/bash
# Round evidence run (one gpurun call): GPU tests, both bench arms, ncu launch list, per-kernel DRAM traffic, one --set full capture.
# Only text / compressed CSV travels back (gpurun_out is capped at 64 MiB).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cat $O/bench_reference.json
echo "== bench b200"; timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "== ncu DRAM bytes of every kernel of one frame"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 250 --csv \
    --log-file $O/dram_one_frame.csv python tools/one_frame.py dragon 1024 1024 256 > $O/one_frame_ncu.log 2>&1
python tools/dram_traffic.py $O/dram_one_frame.csv $O/dram_traffic.json > $O/dram_traffic.txt 2>&1; cat $O/dram_traffic.txt
echo "== ncu --set full of the first launches of each kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(primary|trace|shade)' -c 5 -f -o /tmp/full_first5 \
    python tools/one_frame.py dragon 1024 1024 256 > $O/full_ncu.log 2>&1
python tools/ncu_summary.py /tmp/full_first5.ncu-rep > $O/full_first5.txt 2>&1
ncu -i /tmp/full_first5.ncu-rep --page source --csv --kernel-name regex:k_trace --launch-count 1 2>/dev/null | cut -d, -f1-12 | gzip -9 > $O/k_trace_source.csv.gz
rm -f $O/*.ncu-rep
du -sh $O
