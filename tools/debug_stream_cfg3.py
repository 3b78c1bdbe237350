import faulthandler, os, sys
faulthandler.dump_traceback_later(int(os.environ.get("WD", "25")), exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
pkg = entry.import_package()
import torch
from itnn_b200 import problems
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
p = problems.make_config(name)
def log(*a):
    print(*a, file=sys.stderr, flush=True)
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p)
    log("uploaded", ctx.buckets())
    r, d = ctx.sweep(1, 0.0)
    log("plain sweep ok", r)
    flat = ctx.pack_messages(p.messages)
    tdt = torch.complex128 if flat.dtype.kind == "c" else torch.float64
    pa, pb = torch.empty(flat.size, dtype=tdt).pin_memory(), torch.empty(flat.size, dtype=tdt).pin_memory()
    a, b = pa.numpy(), pb.numpy()
    a[:] = flat
    log("streamed call 1 ...")
    r = ctx.sweep_host(a, b)
    log("streamed call 1 ok", r)
    r = ctx.sweep_host(b, a)
    log("streamed call 2 ok", r)
    x, y = flat.copy(), np.empty_like(flat)
    r = ctx.sweep_host(x, y)
    log("staged ok", r)
