#!/bin/bash
# Round 2, GPU call 17: racecheck after the packet-stack fence; per-launch timeline of ONE of 8 ranks (single arena, serial launches).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/one_frame.py dragon 96 96 8 > $O/sanitizer_racecheck_dragon.log 2>&1; grep -E "RACECHECK SUMMARY" $O/sanitizer_racecheck_dragon.log
timeout 600 python -m pytest tests/test_gpu_traversal.py -q -m gpu -k "packets" 2>&1 | tail -2
B200PT_DUMP_TIMELINE=1 timeout 300 python tools/gpu_rank_breakdown.py 8 > $O/rank8_timeline.log 2>&1; grep -c timeline $O/rank8_timeline.log; grep "timeline" $O/rank8_timeline.log | awk '{print $3, $5, $7}' | tr '\n' ';' | cut -c1-3000; echo; tail -1 $O/rank8_timeline.log
B200PT_DUMP_TIMELINE=1 timeout 300 python tools/gpu_rank_breakdown.py 1 > $O/rank1_timeline.log 2>&1; grep "timeline" $O/rank1_timeline.log | awk '{print $3, $5, $7}' | tr '\n' ';' | cut -c1-3000; echo; tail -1 $O/rank1_timeline.log
