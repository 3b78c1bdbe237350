"""Run a few sweeps of one workload (for ncu): python tools/profile_sweep.py cfg2 [nsweeps] [kernel]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.import_package()
from itnn_b200 import graphs, problems

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if name == "cfg5s":
    p = problems.make_config("cfg5", graph=graphs.named_grid((16, 16)))
else:
    p = problems.make_config(name)
with pkg.BPXContext(0) as ctx:
    problems.upload(ctx, p, kernel or None)
    res, done = ctx.sweep(n, 0.0)
    print(name, "sweeps", done, "residual", res, ctx.buckets())
