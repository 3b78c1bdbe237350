#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
T=$PWD/itensornetworksnext.jl_b200/csrc/libbpx_timing.so
K='cfg5 or sampled_edges or fast_kernels or streamed_io or converges_like'
G=${G:-8}
echo "== tests G=$G"
BPX_SLICED_G=$G timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -4
BPX_LIB=$T BPX_SLICED_G=$G timeout 300 python tools/timing_sliced2.py 96 96 2>&1 | tail -16 | grep -v "prod arr\|prod group\|wait MSG\|epi total" | tee $O/r2h_timing_g$G.txt
BPX_SLICED_G=$G timeout 600 python bench.py --no-others --no-cpu-baseline --no-beliefs --steps 5 > $O/r2h_bench_g$G.json 2> $O/r2h_bench_g$G.err
tail -c 400 $O/r2h_bench_g$G.err
python - <<PY
import json
try:
    d = json.load(open("$O/r2h_bench_g$G.json"))
    print("G=$G", "ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["max_rel_err"], "conv", d["convergence"]["sweeps"], d["convergence"]["ms"], d["clocks"])
except Exception as ex:
    print("failed", ex)
PY
BPX_SLICED_G=$G timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_update_sliced_c16g -s 2 -c 1 -f -o $O/r2h_sliced2_g$G \
  python tools/timing_sliced2.py 64 64 > $O/r2h_ncu_g$G.log 2>&1
ncu -i $O/r2h_sliced2_g$G.ncu-rep --page raw --csv > $O/r2h_sliced2_g$G.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2h_sliced2_g$G.raw.csv $O/r2h_sliced2_g${G}_ncu_summary.csv bp_update_sliced 2>&1 | tail -1
head -20 $O/r2h_sliced2_g${G}_ncu_summary.csv
rm -f $O/r2h_sliced2_g$G.ncu-rep
