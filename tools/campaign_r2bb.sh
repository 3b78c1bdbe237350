#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2bb}
timeout 900 python -m pytest tests/test_zz_gpu_apply.py tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzzzz_padded_dims_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for ov in 1 0; do
for cfg in "64 64 --chi 16" "48 48 --chi 16" "32 32 --chi 8" "64 64 --chi 8" "32 32 --chi 16 --dtype c128"; do
  BPX_APPLY_OVERLAP=$ov timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('overlap $ov  $cfg', round(d['value']), d['ms_per_layer'], d['gates_on_gram_kernel'], d['gates_declined_to_stepwise_kernel'])"
done
done
