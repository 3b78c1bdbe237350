#!/bin/bash
set -u
O=gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --target-processes all --log-file $O/r2af_sanitizer_racecheck_v3.log \
  python -m pytest tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzzzz_bethe_device.py -m gpu -q -p no:cacheprovider -k "v3 or bethe or norm_network_configs or signed_single_layer or zero_edge" > $O/r2af_sanitizer_racecheck_v3_pytest.txt 2>&1
tail -2 $O/r2af_sanitizer_racecheck_v3_pytest.txt; tail -3 $O/r2af_sanitizer_racecheck_v3.log
grep -o "in [a-z_0-9]*\.cuh:[0-9]*" $O/r2af_sanitizer_racecheck_v3.log | sort | uniq -c | sort -rn | head
