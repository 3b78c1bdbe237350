#!/bin/bash
# Round 2, consolidation on one B200: whole GPU suite (gate kernel v3 is the default), smoke, default bench + reference arm,
# gate-layer bench, ncu --set full of the sliced kernel ON THE BENCHED LATTICE (256x256), launch list of the default bench.
set -u
O=gpurun_out
T=${TAG:-r2z}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/${T}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/${T}_smoke.txt
S=$(date +%s); timeout 900 python bench.py > $O/${T}_bench_default_n1.json 2> $O/${T}_bench_default_n1.err; echo "default bench rc=$? wall $(( $(date +%s) - S )) s"
timeout 600 python bench.py --impl reference > $O/${T}_bench_reference_n1.json 2> $O/${T}_bench_reference_n1.err
timeout 900 python bench.py --workload apply --steps 8 --warmup 2 > $O/${T}_bench_apply.json 2> $O/${T}_bench_apply.err; tail -c 300 $O/${T}_bench_apply.err
timeout 600 python tools/bench_apply.py --lattice 32 32 --chi 8 --dtype c128 --layers 8 --oracle-gates 0 > $O/${T}_apply_v3_3232chi8c128.json 2> $O/${T}_apply_v3_3232chi8c128.err
timeout 600 python tools/bench_apply.py --lattice 16 16 --chi 16 --dtype c128 --layers 4 --oracle-gates 0 > $O/${T}_apply_v3_1616chi16c128.json 2> $O/${T}_apply_v3_1616chi16c128.err
timeout 300 python tools/bench_simple_update.py --lattice 32 32 --chi 8 --steps 5 > $O/${T}_simple_update_32x32_chi8.json 2> $O/${T}_simple_update_32x32_chi8.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-others --no-cpu-baseline > $O/${T}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bp_update_sliced_c16g -s 1 -c 1 -f -o $O/${T}_sliced2_256 \
  python bench.py --steps 1 --warmup 1 --no-others --no-cpu-baseline --no-beliefs --no-e2e --no-parity --converge 0 > $O/${T}_ncu_sliced_256.log 2>&1; echo "ncu sliced rc=$?"
ncu -i $O/${T}_sliced2_256.ncu-rep --page raw --csv > $O/${T}_sliced2_256.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${T}_sliced2_256.raw.csv $O/${T}_sliced2_256x256_ncu_summary.csv bp_update_sliced 2>&1 | tail -1
head -12 $O/${T}_sliced2_256x256_ncu_summary.csv
rm -f $O/${T}_sliced2_256.ncu-rep
python - <<PY
import json
d=json.load(open("$O/${T}_bench_default_n1.json")); print({k:d[k] for k in ('value','ms_per_step','beliefs','gpu_launches')}); print(d['cpu_baseline']); print(d['e2e']['value'], d['parity']['max_rel_err'], d['convergence']['sweeps'], d['clocks'])
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'))
a=json.load(open("$O/${T}_bench_apply.json")); print('apply', a['value'], a['ms_per_step'], a['roofline']['frac'], a['cpu_baseline'], a['gates_on_gram_kernel'], a['gates_declined_to_stepwise_kernel'])
for f in ("apply_v3_3232chi8c128","apply_v3_1616chi16c128"):
    a=json.load(open("$O/${T}_"+f+".json")); print(f, a['value'], a['ms_per_layer'], a['gates_on_gram_kernel'], a['gates_declined_to_stepwise_kernel'])
PY
