#!/bin/bash
set -u
O=gpurun_out
for sh in 0 1 2; do
  for cfg in "64 64 --chi 16" "32 32 --chi 8"; do
    n=$(echo $cfg | tr -d ' -' )
    BPX_APPLY_GS_SHIFT=$sh BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice $cfg --layers 2 --warmup 1 --oracle-gates 0 > $O/r2ae_t.json 2> $O/r2ae_t.err
    echo "== shift $sh $cfg"; grep "svd  \|eig 0\|total\|Jacobi" $O/r2ae_t.err | tail -4
    BPX_APPLY_GS_SHIFT=$sh timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 > $O/r2ae_b.json 2>/dev/null
    python -c "import json; d=json.load(open('$O/r2ae_b.json')); print(d['value'], d['ms_per_layer'])"
  done
done
