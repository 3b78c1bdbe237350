#!/bin/bash
set -u
for g in 296 222 148 74; do
  echo "== sides grid $g"
  BPX_APPLY_SIDES_GRID=$g BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 2 --warmup 1 --oracle-gates 0 2>&1 >/dev/null | grep "kernel time\|sides: absorb\|sides: gram" | tail -5
done
