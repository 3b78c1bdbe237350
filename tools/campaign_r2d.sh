#!/bin/bash
# sliced kernel v2 (group-cooperative): parity tests for G = 8 / 4 / v1, then the cfg5 bench per variant
set -u
O=gpurun_out
mkdir -p $O
K='cfg5 or sampled_edges or fast_kernels or streamed_io or converges_like'
for var in "BPX_SLICED_G=8" "BPX_SLICED_G=4" "BPX_SLICED_V1=1"; do
  echo "== $var"
  env $var timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -5
done
for var in "BPX_SLICED_G=8" "BPX_SLICED_G=4" "BPX_SLICED_V1=1"; do
  echo "== bench $var"
  env $var timeout 600 python bench.py --no-others --no-cpu-baseline --no-beliefs --steps 5 > $O/r2d_bench_${var}.json 2> $O/r2d_bench_${var}.err
  tail -c 600 $O/r2d_bench_${var}.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r2d_bench_${var}.json"))
    print("$var", "ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["max_rel_err"], "conv", d["convergence"]["sweeps"], d["convergence"]["ms"])
except Exception as ex:
    print("$var failed", ex)
PY
done
