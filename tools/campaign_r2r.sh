#!/bin/bash
# Gram-path gate kernel (v3): GPU parity tests of the whole gate path, device-side bethe free energy, first timing.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_gpu_apply.py tests/test_zz_apply_mirror.py tests/test_zzz_late_gpu.py tests/test_zzz_resident_state.py tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzzzz_bethe_device.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee $O/r2r_pytest_apply_gpu.txt
for cfg in "32 32 --chi 8" "32 32 --chi 8 --dtype c128" "16 16 --chi 16" "64 64 --chi 16"; do
  n=$(echo $cfg | tr -d ' -' )
  timeout 600 python tools/bench_apply.py --lattice $cfg --layers 8 --oracle-gates 0 > $O/r2r_apply_v3_$n.json 2> $O/r2r_apply_v3_$n.err
  echo "== v3 $cfg"; tail -c 500 $O/r2r_apply_v3_$n.json; echo; tail -c 300 $O/r2r_apply_v3_$n.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2r_launches_apply_v3.csv \
  python tools/bench_apply.py --lattice 64 64 --chi 16 --layers 2 --oracle-gates 0 > $O/r2r_apply_under_ncu.log 2>&1
grep -c bp_apply $O/r2r_launches_apply_v3.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_apply_gates_v3 -c 1 -o $O/r2r_apply_v3_chi16 -f \
  python tools/bench_apply.py --lattice 32 32 --chi 16 --layers 1 --warmup 0 --oracle-gates 0 > $O/r2r_apply_ncu_full.log 2>&1
ncu -i $O/r2r_apply_v3_chi16.ncu-rep --page raw --csv > $O/r2r_apply_v3_chi16.raw.csv 2>/dev/null
ncu -i $O/r2r_apply_v3_chi16.ncu-rep --page source --csv > $O/r2r_apply_v3_chi16.source.csv 2>/dev/null
python tools/ncu_summary.py $O/r2r_apply_v3_chi16.raw.csv $O/r2r_apply_v3_chi16_ncu_summary.csv bp_apply_gates 2>&1 | tail -2
rm -f $O/r2r_apply_v3_chi16.ncu-rep
timeout 300 python tools/bench_simple_update.py --lattice 32 32 --chi 8 --steps 5 > $O/r2r_simple_update_32x32_chi8.json 2> $O/r2r_simple_update_32x32_chi8.err
tail -c 700 $O/r2r_simple_update_32x32_chi8.json
