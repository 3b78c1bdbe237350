#!/bin/bash
# final evidence for the gate path: ncu --set full of the side kernel (column-group-major layout) and the bond kernel, launch list of bench.py --workload apply
set -u
O=gpurun_out
T=r2bq
for k in sides bond; do
  timeout 600 ncu --set full --clock-control none -k regex:bp_apply3_$k -c 1 -o $O/${T}_apply3_$k -f \
    python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 1 --warmup 0 --oracle-gates 0 > $O/${T}_ncu_$k.log 2>&1
  ncu -i $O/${T}_apply3_$k.ncu-rep --page raw --csv > $O/${T}_apply3_$k.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/${T}_apply3_$k.raw.csv $O/${T}_apply3_${k}_ncu_summary.csv bp_apply3 2>&1 | tail -1
  rm -f $O/${T}_apply3_$k.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench_apply.csv python bench.py --workload apply --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_launches.log 2>&1
grep -c "bp_apply3" $O/${T}_launches_bench_apply.csv
