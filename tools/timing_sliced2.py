"""Debug: per-CTA cycle counters of bp_update_sliced_c16g (needs a BPX_SLICED_TIMING build:
   BPX_BUILD_OUT=.../libbpx_timing.so NVCC_EXTRA=-DBPX_SLICED_TIMING python csrc/build.py --force; run with BPX_LIB=that)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.import_package()
from itnn_b200 import graphs, problems
dims = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 64)
p = problems.make_config("cfg5", graph=graphs.named_grid(dims), host_data=False)
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p)
    ctx.sweep(2)
    buf = np.zeros(8 * 32 * 16, dtype=np.int64)
    ctx.lib.bpx_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    ctx.lib.bpx_debug_timing(ctx.h, None, 0)  # allocate (zeroed)
    ctx.sweep(1)
    ctx.lib.bpx_debug_timing(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size)
    t = buf[:148 * 16].reshape(148, 16).astype(float)
    names = ["compute total", "wait FULL S1", "wait FULL S2", "wait FULL S3", "dump", "wait MSG", "prod group_wait", "prod arrive",
             "prod wait DONE", "prod bulk_wait", "prod total", "epi B3 wait", "wait FULL S3 q=0", "epi total"]
    tot = t[:, 0].mean()
    print(f"lattice {dims}, G = {os.environ.get('BPX_SLICED_G', '8')}: compute-warp-0 cycles per CTA {tot:.0f}")
    for i, nm in enumerate(names):
        print(f"  {nm:16s} mean {t[:, i].mean():12.0f}  ({100 * t[:, i].mean() / tot:5.1f} % of compute total)   min {t[:, i].min():12.0f} max {t[:, i].max():12.0f}")
