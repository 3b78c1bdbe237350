#!/bin/bash
# One short GPU call (a few GPU-minutes were left; r1i = first run, r1j = after the staged rewrite): validates the single-layer vertex kernel, re-runs the whole GPU suite,
# the default bench line (sanity after the bench.py / device.py edits), the new HBM-bound workload, and one ncu capture.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r1m_pytest_gpu.txt
echo "pytest done at $(( $(date +%s) - T0 )) s" >> $O/r1m_pytest_gpu.txt
timeout 80 python bench.py --steps 10 --warmup 3 --cpu-seconds 3 > $O/r1m_bench_cfg2_n1.json 2> $O/r1m_bench_cfg2_n1.err
timeout 120 python bench.py --workload ising --steps 20 --warmup 3 --cpu-seconds 5 > $O/r1m_bench_ising_n1.json 2> $O/r1m_bench_ising_n1.err
timeout 60 python -c "import __graft_entry__ as e; e.smoke()" > $O/r1m_smoke.txt 2>&1
echo "all done at $(( $(date +%s) - T0 )) s" >> $O/r1m_pytest_gpu.txt
cat $O/r1m_pytest_gpu.txt
tail -c 600 $O/r1m_bench_ising_n1.err
head -c 900 $O/r1m_bench_ising_n1.json
echo
head -c 400 $O/r1m_bench_cfg2_n1.json
echo
cat $O/r1m_smoke.txt | tail -2
