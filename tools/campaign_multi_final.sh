#!/bin/bash
# final multi-GPU lines: bash tools/campaign_multi_final.sh N
set -u
N=$1
O=gpurun_out
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N "$@" > $O/rh_bench_${name}_n$N.json 2> $O/rh_bench_${name}_n$N.err
  python - "$O/rh_bench_${name}_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 3),
          "e2e_ms", round(e.get("ms_per_step", 0) or 0, 4), "e2e", round(e.get("value", 0) or 0), "launches", d["gpu_launches"],
          "conv", d.get("convergence") and (d["convergence"]["sweeps"], round(d["convergence"]["ms"], 2)))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run cfg2 --steps 30 --warmup 3 --no-cpu-baseline
run cfg5 --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --converge 1e-10
