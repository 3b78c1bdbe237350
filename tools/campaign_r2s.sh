#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2s}
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_gpu_apply.py tests/test_zzzz_apply_large_and_v2_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/${T}_pytest_apply_gpu.txt
for cfg in "32 32 --chi 8" "16 16 --chi 16" "64 64 --chi 16" "32 32 --chi 8 --dtype c128"; do
  n=$(echo $cfg | tr -d ' -' )
  timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 > $O/${T}_apply_v3_$n.json 2> $O/${T}_apply_v3_$n.err
  echo "== v3 $cfg"; python -c "import json; d=json.load(open('$O/${T}_apply_v3_$n.json')); print(d['value'], d['ms_per_layer'], d['gates_on_gram_kernel'], d['gates_declined_to_stepwise_kernel'])"; tail -c 300 $O/${T}_apply_v3_$n.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply_gates_v3 -c 1 -o $O/${T}_apply_v3_chi16 -f \
  python tools/bench_apply.py --lattice 32 32 --chi 16 --layers 1 --warmup 0 --oracle-gates 0 > $O/${T}_apply_ncu_full.log 2>&1
ncu -i $O/${T}_apply_v3_chi16.ncu-rep --page raw --csv > $O/${T}_apply_v3_chi16.raw.csv 2>/dev/null
ncu -i $O/${T}_apply_v3_chi16.ncu-rep --page source --csv > $O/${T}_apply_v3_chi16.source.csv 2>/dev/null
python tools/ncu_summary.py $O/${T}_apply_v3_chi16.raw.csv $O/${T}_apply_v3_chi16_ncu_summary.csv bp_apply_gates 2>&1 | tail -1
rm -f $O/${T}_apply_v3_chi16.ncu-rep

