#!/bin/bash
set -u
O=gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_partition.py -m gpu -q -p no:cacheprovider -k "multi_launch" 2>&1 | tail -25 > $O/r2bf_run_$i.txt
  tail -1 $O/r2bf_run_$i.txt
done
N=2 TAG=r2bf bash tools/campaign_multi_r2.sh
