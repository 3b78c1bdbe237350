#!/bin/bash
set -u
O=gpurun_out
for ct in 5 4 3; do
  for cfg in "64 64 --chi 16" "32 32 --chi 8"; do
    echo "== bond CTAs $ct, $cfg"
    BPX_APPLY_BOND_CTAS=$ct BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice $cfg --layers 2 --warmup 1 --oracle-gates 0 2>&1 >/dev/null | tail -16 | grep -v "^  sides\|^  final"
  done
done
