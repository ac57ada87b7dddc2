#!/bin/bash
# Round 2, GPU call 29: per-kernel ncu counters of matpreview and volumetric-caustic again (k_bin_hits / k_settle were rewritten after the first capture).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
M=$(python tools/ncu_counters.py --metrics)
for cfg in "volumetric-caustic 1024 1024 256" "matpreview 1024 1024 128"; do
  set -- $cfg
  echo "== ncu counters $cfg"
  timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/raw_$1.csv python tools/one_frame.py $cfg > $O/one_frame_$1.log 2>&1
  python tools/ncu_counters.py $O/raw_$1.csv $O/r02_counters_$1_$2x$3x$4.json "$1 $2x$3x$4" 2>&1 | head -12
  rm -f $O/raw_$1.csv
done
