#!/bin/bash
# Vertex-kernel tuning variants on one box: register cap (5 vs 6 resident CTAs) x grid policy (persistent vs 32 CTAs per SM).
set -u
O=gpurun_out
mkdir -p $O
run() {  # name, lib, ctas_per_sm
  BPX_LIB=$2 BPX_VERTEX_CTAS_PER_SM=$3 timeout 60 python bench.py --workload ising --steps 30 --warmup 3 --no-cpu-baseline --no-e2e \
    > $O/r1l_ising_$1.json 2> $O/r1l_ising_$1.err
  python - "$O/r1l_ising_$1.json" "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 5), "launch_ms", round(d["roofline"]["avg_launch_ms"], 5), "frac", round(d["roofline"]["frac"], 4))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
D=$PWD/itensornetworksnext.jl_b200/csrc/libbpx.so
V=$PWD/tools/_variants/libbpx_min6.so
run min5_persistent $D 0
run min6_persistent $V 0
run min5_32persm $D 32
run min6_32persm $V 32
run min5_persistent_again $D 0
