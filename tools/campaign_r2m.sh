#!/bin/bash
set -u
O=gpurun_out
T=$PWD/itensornetworksnext.jl_b200/csrc/libbpx_timing.so
K='cfg5 or sampled_edges or fast_kernels or streamed_io or converges_like or group_cooperative'
for CW in 8 16; do
  echo "== CW=$CW"
  BPX_SLICED_CW=$CW timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -3
  BPX_LIB=$T BPX_SLICED_CW=$CW timeout 300 python tools/timing_sliced2.py 96 96 2>&1 | tail -16 | grep -v "prod arr\|prod group\|wait MSG\|epi total\|epi B3" | tee $O/r2m_timing_cw$CW.txt
  BPX_SLICED_CW=$CW timeout 600 python bench.py --no-others --no-cpu-baseline --no-beliefs --steps 5 > $O/r2m_bench_cw$CW.json 2> $O/r2m_bench_cw$CW.err
  tail -c 300 $O/r2m_bench_cw$CW.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r2m_bench_cw$CW.json"))
    print("CW=$CW", "ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["max_rel_err"], "conv", d["convergence"]["sweeps"], d["convergence"]["ms"], d["clocks"])
except Exception as ex:
    print("failed", ex)
PY
done
