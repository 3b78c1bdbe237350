#!/bin/bash
# Multi-GPU measurement campaign: bash tools/campaign_multi.sh N   (run under gpurun --gpus N)
set -u
N=$1
O=gpurun_out
mkdir -p $O
run() {  # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N "$@" > $O/rg_bench_${name}_n$N.json 2> $O/rg_bench_${name}_n$N.err
  python - "$O/rg_bench_${name}_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], "ms/step", round(d.get("ms_per_step", 0), 4), "value", round(d.get("value", 0)), "frac", round(r.get("frac", 0), 3),
          "e2e_ms", round(e.get("ms_per_step", 0) or 0, 4), "e2e", round(e.get("value", 0) or 0), "conv", d.get("convergence") and (d["convergence"]["sweeps"], round(d["convergence"]["ms"], 2)))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run cfg2 --steps 30 --warmup 3
run reference --impl reference --steps 5 --warmup 1
run cfg5 --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --converge 1e-10
run cfg4 --workload cfg4 --steps 30 --warmup 3 --no-cpu-baseline
