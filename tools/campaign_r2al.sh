#!/bin/bash
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee $O/r2al_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > $O/r2al_bench_default_n1.json 2> $O/r2al_bench_default_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('$O/r2al_bench_default_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['parity']['max_rel_err'], d['convergence']['sweeps'], d['beliefs']['ms'])
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'))"
