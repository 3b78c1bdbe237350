#!/bin/bash
set -u
C=itensornetworksnext.jl_b200/csrc
timeout 600 python -m pytest tests/test_zz_gpu_apply.py tests/test_zzzz_apply_large_and_v2_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for lib in libbpx.so libbpx_b.so; do
  [ -f $C/$lib ] || continue
  for cfg in "64 64 --chi 16" "32 32 --chi 8" "16 16 --chi 16 --dtype c128" "32 32 --chi 8 --dtype c128"; do
    case "$cfg" in *c128*) cts="2 3";; *) cts="4 5";; esac
    for ct in $cts; do
    echo "== $lib $cfg bond CTAs $ct"
    BPX_LIB=$PWD/$C/$lib BPX_APPLY_BOND_CTAS=$ct BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice $cfg --layers 2 --warmup 1 --oracle-gates 0 2>&1 >/dev/null | tail -15 | grep "kernel time\|bond: svd\|bond: eig"
    BPX_LIB=$PWD/$C/$lib BPX_APPLY_BOND_CTAS=$ct timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('   gates/s', round(d['value']), d['ms_per_layer'])"
    done
  done
done
