"""bench.py --single-process --gpus N: the SAME workload driven from ONE process through bpx_create_multi (one context
over N devices; what a single `beliefpropagation()` call of the Julia plugin does).  Prints one JSON line shaped like
bench.py's.  Timing: wall clock around K x (enqueue one sweep on every device + synchronise all devices), after W warm-up
sweeps -- the devices' own kernels gate each other over NVLink, the host only enqueues."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def run(args):
    import bench

    pkg = entry.import_package()
    from itnn_b200 import problems

    n = args.gpus
    p, owner, desc, scaling = bench.build_workload(args.workload, n)
    ctx = pkg.BPXContext(devices=list(range(n)))
    ctx.set_graph(p.ga.src, p.ga.dst, p.ga.slot, p.ga.nv)
    ctx.set_dims(p.dtype, p.mode, p.phys_dim if p.mode == "norm" else None, p.link_dim)
    if owner is not None:
        ctx.set_owner(owner)  # the same strips / slabs as the one-process-per-GPU run
    if p.tensors is None:
        ctx.fill_synthetic(123)
    else:
        ctx.set_site_tensors(p.tensors)
        ctx.set_messages(p.messages)
    for _ in range(args.warmup):
        ctx.sweep_async(1)
    ctx.synchronize()
    ctx.counters(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.sweep_async(1)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    counters = ctx.counters()
    ms = 1e3 * dt / args.steps
    value = p.ga.ne / (ms * 1e-3)
    D = bench.Dist(0, 1)
    # parity: sampled edges from the devices' own inputs (owner lookup inside the library), cut edges included
    class _View:  # what parity_check needs from a per-rank context, served by the multi-device context
        def __init__(self, c):
            self.c = c
            self.msg_off, self.link_dim = c.msg_off, c.link_dim

        def __getattr__(self, k):
            return getattr(self.c, k)

    parity = None if args.no_parity else bench.parity_check(_View(ctx), p, None, D, args.parity_edges)
    conv = None
    if args.converge > 0:
        if p.tensors is None:
            ctx.fill_synthetic(123)
        else:
            ctx.set_messages(p.messages)
        t0 = time.perf_counter()
        res, done = ctx.sweep(200, args.converge)
        tc = time.perf_counter() - t0
        conv = {"tol": args.converge, "sweeps": int(done), "residual": res, "ms": 1e3 * tc, "converged": bool(res < args.converge)}
    line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64" if p.dtype.kind != "c" else "c128",
            "data": "synthetic", "mode": "single process, one bpx_create_multi context over all devices",
            "config": {"workload": desc, "sharding": bench.shard_note(args.workload, n), "updates_per_step": p.ga.ne,
                       "timing": "host wall clock around K sweeps enqueued on every device + one synchronise (L2 not flushed)"},
            "gpu_launches": int(counters["launches"]), "parity": parity, "convergence": conv}
    print(json.dumps(line), flush=True)
    ctx.close()
    if parity is not None and not parity["ok"]:
        sys.exit(3)
