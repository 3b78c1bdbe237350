#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2bg}
for k in sides final; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply3_$k -c 1 -o $O/${T}_apply3_$k -f \
    python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 1 --warmup 0 --oracle-gates 0 > $O/${T}_ncu_$k.log 2>&1
  ncu -i $O/${T}_apply3_$k.ncu-rep --page raw --csv > $O/${T}_apply3_$k.raw.csv 2>/dev/null
  ncu -i $O/${T}_apply3_$k.ncu-rep --page source --csv > $O/${T}_apply3_$k.source.csv 2>/dev/null
  python tools/ncu_summary.py $O/${T}_apply3_$k.raw.csv $O/${T}_apply3_${k}_ncu_summary.csv bp_apply3 2>&1 | tail -1
  rm -f $O/${T}_apply3_$k.ncu-rep
done
