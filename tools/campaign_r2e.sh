#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
T=$PWD/itensornetworksnext.jl_b200/csrc/libbpx_timing.so
for G in 8 4; do
  BPX_LIB=$T BPX_SLICED_G=$G timeout 300 python tools/timing_sliced2.py 96 96 2>&1 | tail -16 | tee $O/r2e_timing_g$G.txt
done
for G in 8 4; do
  BPX_SLICED_G=$G timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_update_sliced_c16g -s 2 -c 1 -f -o $O/r2e_sliced2_g$G \
    python tools/timing_sliced2.py 48 48 > $O/r2e_ncu_g$G.log 2>&1
  ncu -i $O/r2e_sliced2_g$G.ncu-rep --page raw --csv > $O/r2e_sliced2_g$G.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r2e_sliced2_g$G.raw.csv $O/r2e_sliced2_g${G}_ncu_summary.csv bp_update_sliced 2>&1 | tail -1
done
