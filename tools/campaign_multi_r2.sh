#!/bin/bash
# N GPUs of one box: the multi-GPU tests (N = 2 only) and the driver-style default bench at N
set -u
O=gpurun_out
N=${N:-2}
T=${TAG:-r2zz}
mkdir -p $O
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee $O/${T}_pytest_gpu_2gpus.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/${T}_bench_cfg5_n$N.json 2> $O/${T}_bench_cfg5_n$N.err; echo "bench N=$N rc=$?"; tail -c 400 $O/${T}_bench_cfg5_n$N.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench_cfg5_n$N.json")); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches')}); print(d['e2e']['value'], d['parity'], d['convergence']['sweeps'], d['convergence']['ms'], d['clocks'])
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('parity',{}).get('max_rel_err') if isinstance(v.get('parity'),dict) else None)
PY
