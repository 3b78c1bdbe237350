#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_zzzzz_padded_dims_gpu.py tests/test_gpu_parity.py -k "padded or pad or ragged or multi_device" -m gpu -q -p no:cacheprovider 2>&1 | tail -30 | tee $O/r2ag_pytest_padded.txt
timeout 900 python -m pytest tests/test_zz_gpu_apply.py tests/test_zzzz_apply_large_and_v2_gpu.py tests/test_zzz_resident_state.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python - <<'PY'
# chi = 12 PEPS, 48x48: padded (sliced kernel) vs BPX_NO_PAD (generic kernel)
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as entry
pkg = entry.import_package()
from itnn_b200 import graphs, problems
import torch
out = {}
for tag, env in (("padded", None), ("generic", "1")):
    if env: os.environ["BPX_NO_PAD"] = env
    else: os.environ.pop("BPX_NO_PAD", None)
    q = problems.synthetic_peps(graphs.named_grid((48, 48)), 12, 2, np.float64, host_data=False)
    with pkg.BPXContext(0) as ctx:
        problems.upload(ctx, q)
        ctx.sweep(2, 0.0, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5 if env is None else 2
        res, _ = ctx.sweep(n, 0.0, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        out[tag] = {"ms_per_sweep": dt * 1e3, "updates_per_s": q.ga.ne / dt, "residual": res,
                    "buckets": [(b["degree"], b["chi"], b["kernel"]) for b in ctx.buckets()]}
print(json.dumps(out))
open("gpurun_out/r2ag_chi12_48x48.json", "w").write(json.dumps(out))
PY
