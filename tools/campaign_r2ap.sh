#!/bin/bash
# A/B of library builds (Jacobi variants): gate throughput per build
set -u
O=gpurun_out
C=itensornetworksnext.jl_b200/csrc
for lib in libbpx.so libbpx_b.so libbpx_c.so libbpx_d.so; do
  [ -f $C/$lib ] || continue
  for cfg in "64 64 --chi 16" "32 32 --chi 8" "16 16 --chi 16 --dtype c128"; do
    case "$cfg" in *c128*) ct=2;; *) ct=4;; esac
    BPX_LIB=$PWD/$C/$lib BPX_APPLY_BOND_CTAS=$ct timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$lib', '$cfg', round(d['value']), d['ms_per_layer'])"
  done
done
