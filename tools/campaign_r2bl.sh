#!/bin/bash
# sanitizer passes over the three gate kernels (Gram / final passes with TMA + mbarrier, joint Jacobi) incl. the fallback hand-over
set -u
O=gpurun_out
T=${TAG:-r2bl}
for tool in memcheck synccheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --target-processes all --log-file $O/${T}_sanitizer_$tool.log \
    python -m pytest tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zz_gpu_apply.py -m gpu -q -p no:cacheprovider -k "v3 or layer or padded or two_site" > $O/${T}_sanitizer_${tool}_pytest.txt 2>&1
  tail -2 $O/${T}_sanitizer_${tool}_pytest.txt
  tail -3 $O/${T}_sanitizer_$tool.log
  grep -o "in [a-z_0-9]*\.cuh:[0-9]*" $O/${T}_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -8
done
