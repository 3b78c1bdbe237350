"""Debug: per-phase clock64 stamps of CTA 0 (needs a BPX_ONCHIP_TIMING build: NVCC_EXTRA=-DBPX_ONCHIP_TIMING)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.import_package()
from itnn_b200 import problems
p = problems.make_config("cfg2")
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p)
    ctx.sweep(2)
    buf = np.zeros(8 * 32 * 16, dtype=np.int64)
    ctx.lib.bpx_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    ctx.lib.bpx_debug_timing(ctx.h, None, 0)  # allocate
    ctx.sweep(1)
    ctx.lib.bpx_debug_timing(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size)
    t = buf.reshape(8, 32, 16)
    names = ["top", "-", "mbar_wait", "frags", "pair", "PHASEwait", "closeA", "closeB", "publish"]
    for it in range(7):
        w = t[it, :16, :9]; w[:, 1] = w[:, 0]
        t0 = w[:, 0].min()
        print(f"item {it}: start {t0 - t[0,:16,0].min()}")
        for i in range(2, 9):
            d = w[:, i] - w[:, i - 1]
            print(f"   {names[i]:14s} mean {d.mean():8.0f}  min {d.min():6d}  max {d.max():6d}   (warp-end spread {w[:, i].max() - w[:, i].min()})")
        print(f"   total {w[:, 8].max() - t0}")
