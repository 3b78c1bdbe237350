#!/bin/bash
# Round 2, session 2, first call: the driver's default commands end to end (wall clock, cpu_baseline, reference arm), the
# launch list of the default bench, and the first timing of the gate-application kernels (v1 / v2).
set -u
O=gpurun_out
mkdir -p $O
S=$(date +%s)
timeout 900 python bench.py > $O/r2o_bench_default_n1.json 2> $O/r2o_bench_default_n1.err; echo "default bench rc=$? wall $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 600 python bench.py --impl reference > $O/r2o_bench_reference_n1.json 2> $O/r2o_bench_reference_n1.err; echo "reference arm rc=$? wall $(( $(date +%s) - S )) s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2o_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-others --no-cpu-baseline > $O/r2o_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
# gate application: first timing
timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 8 > $O/r2o_apply_32x32_chi8.json 2> $O/r2o_apply_32x32_chi8.err
timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --dtype c128 --layers 8 > $O/r2o_apply_32x32_chi8_c128.json 2> $O/r2o_apply_32x32_chi8_c128.err
timeout 600 python tools/bench_apply.py --lattice 16 16 --chi 16 --layers 4 --oracle-gates 2 > $O/r2o_apply_16x16_chi16.json 2> $O/r2o_apply_16x16_chi16.err
BPX_APPLY_V2=1 timeout 300 python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 8 --oracle-gates 0 > $O/r2o_apply_v2_32x32_chi8.json 2> $O/r2o_apply_v2_32x32_chi8.err
BPX_APPLY_V2=1 timeout 600 python tools/bench_apply.py --lattice 16 16 --chi 16 --layers 4 --oracle-gates 0 > $O/r2o_apply_v2_16x16_chi16.json 2> $O/r2o_apply_v2_16x16_chi16.err
timeout 300 python tools/bench_simple_update.py --lattice 32 32 --chi 8 --steps 5 > $O/r2o_simple_update_32x32_chi8.json 2> $O/r2o_simple_update_32x32_chi8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2o_launches_apply.csv \
  python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 2 --oracle-gates 0 > $O/r2o_apply_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply_gates -c 1 -o $O/r2o_apply_gates_chi8 -f \
  python tools/bench_apply.py --lattice 32 32 --chi 8 --layers 1 --warmup 0 --oracle-gates 0 > $O/r2o_apply_ncu_full.log 2>&1
ncu -i $O/r2o_apply_gates_chi8.ncu-rep --page raw --csv > $O/r2o_apply_gates_chi8.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2o_apply_gates_chi8.raw.csv $O/r2o_apply_gates_chi8_ncu_summary.csv bp_apply_gates 2>&1 | tail -2
rm -f $O/r2o_apply_gates_chi8.ncu-rep
for f in r2o_apply_32x32_chi8 r2o_apply_32x32_chi8_c128 r2o_apply_16x16_chi16 r2o_apply_v2_32x32_chi8 r2o_apply_v2_16x16_chi16 r2o_simple_update_32x32_chi8; do
  echo "== $f"; tail -c 700 $O/$f.json; echo; tail -c 300 $O/$f.err; done
tail -c 400 $O/r2o_bench_default_n1.err
tail -c 500 $O/r2o_bench_reference_n1.json
