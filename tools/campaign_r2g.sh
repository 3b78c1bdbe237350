#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for G in 8; do
  BPX_SLICED_G=$G timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_update_sliced_c16g -s 2 -c 1 -f -o $O/r2g_sliced2_g$G \
    python tools/timing_sliced2.py 64 64 > $O/r2g_ncu_g$G.log 2>&1
  ncu -i $O/r2g_sliced2_g$G.ncu-rep --page raw --csv > $O/r2g_sliced2_g$G.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r2g_sliced2_g$G.raw.csv $O/r2g_sliced2_g${G}_ncu_summary.csv bp_update_sliced 2>&1 | tail -1
  ncu -i $O/r2g_sliced2_g$G.ncu-rep --page source --csv > $O/r2g_sliced2_g$G.source.csv 2>/dev/null
done
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
