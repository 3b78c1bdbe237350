#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2at}
timeout 900 python -m pytest tests/test_zz_gpu_apply.py tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzzzz_padded_dims_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 | tee $O/${T}_pytest_apply_gpu.txt
for cfg in "64 64 --chi 16" "32 32 --chi 8"; do
  n=$(echo $cfg | tr -d ' -' )
  BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice $cfg --layers 2 --warmup 1 --oracle-gates 0 > $O/${T}_timing_$n.json 2> $O/${T}_timing_$n.err
  echo "== timing $cfg"; tail -16 $O/${T}_timing_$n.err
done
for cfg in "16 16 --chi 16" "64 64 --chi 16" "32 32 --chi 8" "16 16 --chi 16 --dtype c128" "32 32 --chi 8 --dtype c128"; do
  n=$(echo $cfg | tr -d ' -' )
  timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 > $O/${T}_apply_v3_${n}.json 2> $O/${T}_apply_v3_${n}.err
  python -c "import json; d=json.load(open('$O/${T}_apply_v3_${n}.json')); print('$cfg', d['value'], d['ms_per_layer'], d['gates_on_gram_kernel'], d['gates_declined_to_stepwise_kernel'])"; tail -c 300 $O/${T}_apply_v3_${n}.err
done
