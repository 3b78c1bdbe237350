#!/bin/bash
# Single-GPU measurement campaign (run under gpurun); results land in gpurun_out/ and are copied into profiles/ by hand.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/rg_pytest_gpu.txt
timeout 300 python bench.py --steps 30 --warmup 3 > $O/rg_bench_cfg2_n1.json 2> $O/rg_bench_cfg2_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/rg_bench_reference_n1.json 2> $O/rg_bench_reference_n1.err
for w in cfg1 cfg3 cfg4; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --converge 1e-10 > $O/rg_bench_${w}_n1.json 2> $O/rg_bench_${w}_n1.err
done
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --converge 1e-10 --no-cpu-baseline --no-e2e > $O/rg_bench_cfg2_converge_n1.json 2>/dev/null
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --converge 1e-10 > $O/rg_bench_cfg5_n1.json 2> $O/rg_bench_cfg5_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/rg_launches_cfg2_bench.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/rg_bench_under_ncu.log 2>&1
cat $O/rg_pytest_gpu.txt
for f in $O/rg_bench_*_n1.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], "ms/step", round(d.get("ms_per_step", 0), 4), "value", round(d.get("value", 0)), "frac", round(r.get("frac", 0), 3),
          "e2e_ms", round(e.get("ms_per_step", 0) or 0, 4), "conv", d.get("convergence") and (d["convergence"]["sweeps"], round(d["convergence"]["ms"], 2)))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
