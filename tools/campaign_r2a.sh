#!/bin/bash
# Round 2, first GPU call: whole GPU suite (incl. the tests round 1 never ran on a GPU), FP64 peak, sanitizer passes.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -40 > $O/r2a_pytest_gpu.txt
timeout 120 python tools/fp64_peak.py $O/r2a_fp64_peak.json > /dev/null 2> $O/r2a_fp64_peak.err
K='cfg1_4x4 or cfg3_heavy or cfg4_cubic or cfg5_bucket or real_slice_kernel or ragged_link'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --log-file $O/r2a_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$K" > $O/r2a_sanitizer_${tool}_pytest.txt 2>&1
  tail -3 $O/r2a_sanitizer_${tool}_pytest.txt
  tail -5 $O/r2a_sanitizer_$tool.log
done
cat $O/r2a_pytest_gpu.txt
cat $O/r2a_fp64_peak.json
