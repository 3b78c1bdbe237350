#!/bin/bash
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_zzzzz_padded_dims_gpu.py tests/test_zzzzz_multi_device.py tests/test_partition.py tests/test_zzzz_partitioned_apply_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee $O/r2ai_pytest_2gpus.txt
