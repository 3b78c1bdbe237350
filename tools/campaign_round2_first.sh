#!/bin/bash
# First GPU call of round 2 (one B200, run under gpurun, ~6 min of box time): everything that round 1 could not
# measure after its GPU budget ended.  Results land in gpurun_out/r2a_* and are copied into profiles/ by hand.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/campaign_round2_first.sh'
set -u
O=gpurun_out
mkdir -p $O
# 1. the whole GPU suite; the files written after round 1's budget ended sort last (-x reaches the green ones first)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2a_pytest_gpu.txt
# 2. headline bench + reference arm (unchanged path: regression check against profiles/r1m_bench_cfg2_n1.json)
timeout 300 python bench.py --steps 30 --warmup 3 > $O/r2a_bench_cfg2_n1.json 2> $O/r2a_bench_cfg2_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2a_bench_reference_n1.json 2> $O/r2a_bench_reference_n1.err
# 3. first timing of the gate-application kernel: cfg2-, cfg4- and cfg5-shaped layers
timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 8 > $O/r2a_apply_32x32_chi8.json 2> $O/r2a_apply_32x32_chi8.err
timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --dtype c128 --layers 8 > $O/r2a_apply_32x32_chi8_c128.json 2> $O/r2a_apply_32x32_chi8_c128.err
timeout 600 python tools/bench_apply.py --lattice 16 16 --chi 16 --layers 4 --oracle-gates 2 > $O/r2a_apply_16x16_chi16.json 2> $O/r2a_apply_16x16_chi16.err
# 3b. the opt-in version 2 of the kernel (bpx_apply2.cuh), same workloads
BPX_APPLY_V2=1 timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 8 --oracle-gates 0 > $O/r2a_apply_v2_32x32_chi8.json 2> $O/r2a_apply_v2_32x32_chi8.err
BPX_APPLY_V2=1 timeout 600 python tools/bench_apply.py --lattice 16 16 --chi 16 --layers 4 --oracle-gates 0 > $O/r2a_apply_v2_16x16_chi16.json 2> $O/r2a_apply_v2_16x16_chi16.err
# 3c. a whole simple-update step (gate layers + BP sweeps + bond energies) on one context
timeout 300 python tools/bench_simple_update.py --lattice 32 32 --chi 8 --steps 5 > $O/r2a_simple_update_32x32_chi8.json 2> $O/r2a_simple_update_32x32_chi8.err
# 4. launch list + one full ncu capture of bp_apply_gates on the cfg2-shaped layer
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2a_launches_apply.csv \
  python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 2 --oracle-gates 0 > $O/r2a_apply_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply_gates -c 1 -o $O/r2a_apply_gates_chi8 -f \
  python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 1 --warmup 0 --oracle-gates 0 > $O/r2a_apply_ncu_full.log 2>&1
ncu -i $O/r2a_apply_gates_chi8.ncu-rep --page raw --csv > $O/r2a_apply_gates_chi8.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2a_apply_gates_chi8.raw.csv $O/r2a_apply_gates_chi8_ncu_summary.csv bp_apply_gates 2>&1 | tail -2
cat $O/r2a_pytest_gpu.txt
tail -c 600 $O/r2a_apply_32x32_chi8.json; echo
tail -c 600 $O/r2a_apply_16x16_chi16.json; echo
tail -c 400 $O/r2a_apply_v2_32x32_chi8.json; echo
tail -c 400 $O/r2a_apply_v2_16x16_chi16.json; echo
