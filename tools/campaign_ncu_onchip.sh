#!/bin/bash
# source-level ncu dumps (SASS view incl. shared-memory wavefront counters) of the on-chip kernels on cfg2 / cfg4 / cfg3
set -u
O=gpurun_out
T=${TAG:-r2n}
for w in "cfg2 bp_update_onchip_c8" "cfg4 bp_update_onchip_c16" "cfg3 bp_update_onchip_c16x"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o $O/${T}_$1 \
    python bench.py --workload $1 --steps 2 --warmup 2 --no-others --no-cpu-baseline --no-beliefs --no-e2e --no-parity --converge 0 > $O/${T}_ncu_$1.log 2>&1
  ncu -i $O/${T}_$1.ncu-rep --page raw --csv > $O/${T}_$1.raw.csv 2>/dev/null
  ncu -i $O/${T}_$1.ncu-rep --page source --csv > $O/${T}_$1.source.csv 2>/dev/null
  python tools/ncu_summary.py $O/${T}_$1.raw.csv $O/${T}_${1}_ncu_summary.csv $2 2>&1 | tail -1
  rm -f $O/${T}_$1.ncu-rep
  grep -E "gpu__time|bank_conflicts|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" $O/${T}_${1}_ncu_summary.csv
done
