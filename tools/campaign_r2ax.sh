#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2ax}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee $O/${T}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > $O/${T}_bench_default_n1.json 2> $O/${T}_bench_default_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('$O/${T}_bench_default_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['parity']['max_rel_err'], d['convergence']['sweeps'], d['beliefs']['ms'])
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'))"
timeout 600 python bench.py --workload apply > $O/${T}_bench_apply.json 2> $O/${T}_bench_apply.err; echo "apply rc=$?"; python -c "
import json
d=json.load(open('$O/${T}_bench_apply.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d.get('cpu_baseline'))"
for cfg in "16 16 --chi 16" "64 64 --chi 16" "32 32 --chi 8" "16 16 --chi 16 --dtype c128" "32 32 --chi 8 --dtype c128"; do
  n=$(echo $cfg | tr -d ' -' )
  timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 > $O/${T}_apply_v3_${n}.json 2> $O/${T}_apply_v3_${n}.err
  python -c "import json; d=json.load(open('$O/${T}_apply_v3_${n}.json')); print('$cfg', d['value'], d['ms_per_layer'], d['gates_on_gram_kernel'], d['gates_declined_to_stepwise_kernel'])"
done
BPX_APPLY_TIMING=1 timeout 600 python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 2 --warmup 1 --oracle-gates 0 > /dev/null 2> $O/${T}_timing_6464chi16.err; tail -16 $O/${T}_timing_6464chi16.err
timeout 600 python tools/bench_simple_update.py > $O/${T}_simple_update.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/${T}_simple_update.json')); print({k: d[k] for k in list(d)[:12]})"
