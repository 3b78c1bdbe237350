#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
T=$PWD/itensornetworksnext.jl_b200/csrc/libbpx_timing.so
K='cfg5 or sampled_edges or fast_kernels or streamed_io or converges_like'
for G in 8 4; do
  echo "== tests G=$G"
  BPX_SLICED_G=$G timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -4
  BPX_LIB=$T BPX_SLICED_G=$G timeout 300 python tools/timing_sliced2.py 96 96 2>&1 | tail -16 | tee $O/r2f_timing_g$G.txt
done
for var in "BPX_SLICED_G=8" "BPX_SLICED_G=4"; do
  env $var timeout 600 python bench.py --no-others --no-cpu-baseline --no-beliefs --steps 5 > $O/r2f_bench_${var}.json 2> $O/r2f_bench_${var}.err
  tail -c 400 $O/r2f_bench_${var}.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r2f_bench_${var}.json"))
    print("$var", "ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["max_rel_err"], "conv", d["convergence"]["sweeps"], d["convergence"]["ms"], d["clocks"])
except Exception as ex:
    print("$var failed", ex)
PY
done
