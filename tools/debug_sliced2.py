"""Debug: sliced v2 against v1 sweep by sweep on a cfg5-shaped lattice; which edges differ?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.import_package()
from itnn_b200 import graphs, problems
dims = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 64)
p = problems.make_config("cfg5", graph=graphs.named_grid(dims), host_data=False)
def run(env):
    for k, v in env.items():
        os.environ[k] = v
    out = []
    with pkg.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        for s in range(6):
            res, _ = ctx.sweep(1)
            out.append((res, ctx.get_messages_flat()))
    for k in env:
        os.environ.pop(k)
    return out
a = run({"BPX_SLICED_V1": "1"})
b = run({})
src = np.asarray(p.ga.src); deg = np.diff(np.asarray(p.ga.row_ptr))
for s, ((ra, ma), (rb, mb)) in enumerate(zip(a, b)):
    d = np.abs(ma - mb).reshape(p.ga.ne, 256).max(axis=1)
    bad = np.nonzero(d > 1e-12)[0]
    print(f"sweep {s}: residual v1 {ra:.3e} v2 {rb:.3e}; edges differing {len(bad)} of {p.ga.ne}; max diff {d.max():.3e}")
    if len(bad):
        vs = np.unique(src[bad])
        print("   first bad edges", bad[:12], "vertices", vs[:12], "degrees", deg[vs[:12]], "n bad vertices", len(vs))
        break
