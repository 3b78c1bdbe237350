"""Debug: clock64 stamps of CTA 0's first (degree-3) item in the complex chi=16 kernel
(needs a BPX_ONCHIP_TIMING build: NVCC_EXTRA=-DBPX_ONCHIP_TIMING python itensornetworksnext.jl_b200/csrc/build.py --force)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.import_package()
from itnn_b200 import problems
p = problems.make_config("cfg3")
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p)
    ctx.sweep(2)
    buf = np.zeros(8 * 32 * 16, dtype=np.int64)
    ctx.lib.bpx_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    ctx.lib.bpx_debug_timing(ctx.h, None, 0)  # allocate
    ctx.sweep(1)
    ctx.lib.bpx_debug_timing(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size)
    t = buf.reshape(-1, 16)[:8]
    names = ["start", "frags", "s0 wait", "s0 absorb", "s0 bar", "s0 close", "s0 bar2", "s1 wait", "s1 absorb", "s1 bar", "s1 close", "s1 bar2",
             "partial+bar", "epilogue"]
    t0 = t[:, 0].min()
    prev = t[:, 0]
    for i in range(1, 14):
        d = t[:, i] - prev
        print(f"{names[i]:12s} mean {d.mean():8.0f} min {d.min():7d} max {d.max():7d}   at {t[:, i].max() - t0:7d}")
        prev = t[:, i]
    g = buf[2048:2048 + 4 * 148].reshape(148, 4)
    t0 = g[:, 0].min()
    print("per-CTA (globaltimer ns): start offset, duration; clock64 duration")
    for c in range(0, 148, 4):
        print(" ".join(f"[{c+i:3d}: +{g[c+i,0]-t0:5d} {g[c+i,1]-g[c+i,0]:6d}ns {g[c+i,3]-g[c+i,2]:6d}clk]" for i in range(4)))
    print("kernel span ns:", g[:, 1].max() - t0)
