#!/bin/bash
# Round 2, GPU call 6: ncu --set full with SASS source page of the big k_trace launch (depth 1) and of k_primary (packets), Dragon.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(primary|trace)' -c 3 -f -o /tmp/full_r6 \
    python tools/one_frame.py dragon 1024 1024 256 > $O/full_r6.log 2>&1
tail -3 $O/full_r6.log
python tools/ncu_summary.py /tmp/full_r6.ncu-rep > $O/r02_full_first3.txt 2>&1; grep -E "^---|time_duration|inst_executed.sum|issue_active|thread_inst_executed_per|long_scoreboard|stalled_wait|branch_resolving" $O/r02_full_first3.txt
python tools/ncu_source.py /tmp/full_r6.ncu-rep regex:k_trace $O/r02_k_trace_source.csv.gz 0
python tools/ncu_source.py /tmp/full_r6.ncu-rep regex:k_primary $O/r02_k_primary_source.csv.gz 0
du -sh $O
