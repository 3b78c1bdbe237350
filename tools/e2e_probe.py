"""Time bpx_sweep_host (pinned buffers) per call: python tools/e2e_probe.py [workload] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.import_package()
import torch
from itnn_b200 import problems
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
p = problems.make_config(name)
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p)
    flat = ctx.pack_messages(p.messages)
    tdt = torch.float64 if flat.dtype.kind != "c" else torch.complex128
    pa, pb = torch.empty(flat.size, dtype=tdt).pin_memory(), torch.empty(flat.size, dtype=tdt).pin_memory()
    a, b = pa.numpy(), pb.numpy()
    a[:] = flat
    for pinned in (True, False):
        x, y = (a, b) if pinned else (flat.copy(), np.empty_like(flat))
        for _ in range(10):
            ctx.sweep_host(x, y)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.sweep_host(x, y)
        dt = (time.perf_counter() - t0) / steps
        print(f"{name} pinned={pinned} chunks={os.environ.get('BPX_IO_CHUNKS', 'auto')}: {dt * 1e6:.1f} us per call")
