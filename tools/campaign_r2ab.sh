#!/bin/bash
set -u
O=gpurun_out
T=${TAG:-r2ab}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_golden.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $O/${T}_pytest.txt
for w in cfg4 cfg2; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-others --no-cpu-baseline > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  python -c "import json; d=json.load(open('$O/${T}_bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['max_rel_err'], d['e2e']['value'])"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
python -c "import json; d=json.load(open('$O/${T}_bench_cfg5.json')); print('cfg5', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['max_rel_err'], [b['ms_per_launch'] for b in d['buckets']])"
timeout 600 ncu --set full --clock-control none -k regex:bp_update_onchip_c16 -s 2 -c 1 -f -o $O/${T}_cfg4 \
  python bench.py --workload cfg4 --steps 2 --warmup 2 --no-others --no-cpu-baseline --no-beliefs --no-e2e --no-parity --converge 0 > $O/${T}_ncu_cfg4.log 2>&1
ncu -i $O/${T}_cfg4.ncu-rep --page raw --csv > $O/${T}_cfg4.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${T}_cfg4.raw.csv $O/${T}_cfg4_ncu_summary.csv bp_update_onchip_c16 2>&1 | tail -1
grep -E "gpu__time|bank_conflicts|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" $O/${T}_cfg4_ncu_summary.csv
rm -f $O/${T}_cfg4.ncu-rep
