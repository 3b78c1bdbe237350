#!/bin/bash
set -u
O=gpurun_out
for var in "X=0" "BPX_SLICED_FLAGS=1" "BPX_SLICED_FLAGS=2" "BPX_SLICED_FLAGS=3" "BPX_TMAP_NOPROMO=1" "BPX_SLICED_G=4"; do
  env $var timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:bp_update_sliced_c16g -s 2 -c 1 --csv \
    python tools/timing_sliced2.py 64 64 2>/dev/null | grep -E "dram__|gpu__time|lts__" | awk -F'","' -v v="$var" '{print v, $(NF-2), $(NF-1), $NF}'
done
