#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_zzzzz_multi_device.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee $O/r2j_pytest_multi.txt
timeout 600 python bench.py --gpus 2 --single-process --steps 5 > $O/r2j_bench_sp_n2.json 2> $O/r2j_bench_sp_n2.err; tail -c 600 $O/r2j_bench_sp_n2.err
python -c "
import json
d=json.load(open('$O/r2j_bench_sp_n2.json')); print({k:d[k] for k in ('value','ms_per_step','parity','convergence','gpu_launches')})"
