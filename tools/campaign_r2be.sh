#!/bin/bash
set -u
O=gpurun_out
for i in 1 2 3 4 5 6 7 8 9 10; do
  timeout 300 python -m pytest tests/test_partition.py -m gpu -q -p no:cacheprovider -k "multi_launch" 2>&1 | tail -25 > $O/r2be_run_$i.txt
  tail -1 $O/r2be_run_$i.txt
done
grep -l "failed" $O/r2be_run_*.txt | head -3
