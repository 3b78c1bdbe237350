#!/bin/bash
set -u
for v in "BOND 592" "BOND 444" "BOND 296" "BOND 148" "FINAL 296" "FINAL 148" "SIDES 296" "SIDES 148"; do
  set -- $v
  echo "== $1 grid $2"
  env BPX_APPLY_${1}_GRID=$2 BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 2 --warmup 1 --oracle-gates 0 2>&1 >/dev/null | grep "kernel time" | tail -1
done
