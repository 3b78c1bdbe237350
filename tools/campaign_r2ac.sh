#!/bin/bash
# sanitizer passes over the round-2 additions: gate kernel v3 (incl. its fallback hand-over), device-side bethe free energy,
# the c16 kernel's new partial-tile layout
set -u
O=gpurun_out
T=${TAG:-r2ac}
K='v3 or bethe or norm_network_configs or signed_single_layer or zero_edge or cfg4_cubic or cfg5_bucket'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --log-file $O/${T}_sanitizer_$tool.log \
    python -m pytest tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzzzz_bethe_device.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$K" > $O/${T}_sanitizer_${tool}_pytest.txt 2>&1
  tail -2 $O/${T}_sanitizer_${tool}_pytest.txt
  tail -3 $O/${T}_sanitizer_$tool.log
done
