#!/bin/bash
# one ncu --set full capture of bp_apply_gates_v3 on a chi = 16 layer (TAG names the outputs)
set -u
O=gpurun_out
T=${TAG:-r2y}
L=${LATTICE:-"32 32"}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply_gates_v3 -c 1 -o $O/${T}_apply_v3_chi16 -f \
  python tools/bench_apply.py --lattice $L --chi 16 --layers 1 --warmup 0 --oracle-gates 0 > $O/${T}_apply_ncu_full.log 2>&1
ncu -i $O/${T}_apply_v3_chi16.ncu-rep --page raw --csv > $O/${T}_apply_v3_chi16.raw.csv 2>/dev/null
ncu -i $O/${T}_apply_v3_chi16.ncu-rep --page source --csv > $O/${T}_apply_v3_chi16.source.csv 2>/dev/null
python tools/ncu_summary.py $O/${T}_apply_v3_chi16.raw.csv $O/${T}_apply_v3_chi16_ncu_summary.csv bp_apply_gates 2>&1 | tail -1
rm -f $O/${T}_apply_v3_chi16.ncu-rep
