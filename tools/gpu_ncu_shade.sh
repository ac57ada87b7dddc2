#!/bin/bash
# ncu --set full of LATER k_shade launches (bounce 3 and 4) on the shade-bound BASELINE scenes; only CSV exports travel back.
cd "$(dirname "$0")/.."
O=gpurun_out
for s in matpreview volumetric-caustic; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 2 -c 2 -f -o /tmp/shade_$s \
      python tools/one_frame.py $s 1024 1024 16 > $O/shade_$s.log 2>&1
  python tools/ncu_summary.py /tmp/shade_$s.ncu-rep > $O/shade_later_$s.txt 2>&1
  ncu -i /tmp/shade_$s.ncu-rep --page source --csv --launch-count 1 2>/dev/null | cut -d, -f1-12 | gzip -9 > $O/shade_later_${s}_source.csv.gz
done
du -sh $O; ls -la $O | tail -8
