#!/bin/bash
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/r2k_pytest_gpu.txt
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/r2k_bench_cfg5_n1.json 2> $O/r2k_bench_cfg5_n1.err; tail -c 300 $O/r2k_bench_cfg5_n1.err
python -c "
import json
d=json.load(open('$O/r2k_bench_cfg5_n1.json')); print({k:d[k] for k in ('value','ms_per_step','beliefs','convergence')}); print(d['parity']['max_rel_err'], d['e2e']['ms_per_step'])
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('convergence'), v.get('e2e',{}).get('value') if isinstance(v.get('e2e'),dict) else None, v.get('parity',{}).get('max_rel_err') if isinstance(v.get('parity'),dict) else v)"
