"""Throughput of the gate-application path: layers of vertex-disjoint two-site gates
on a square-lattice PEPS through the C ABI (bpx_apply_two_site_gates), next to the numpy apply oracle on a sample.

  python tools/bench_apply.py [--lattice 32 32] [--chi 8] [--dtype f64|c128] [--layers 5] [--oracle-gates 8]

Prints ONE JSON line (gates/s of whole calls: operator upload + work-space allocation + kernel + sync -- the call is
synchronous; the per-launch device time is in the ncu launch list of the same command).  Not part of bench.py's
contract (that measures BASELINE.json's metric, message updates/s); written for the round-2 measurement of §4.9.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=2, default=[32, 32])
    ap.add_argument("--chi", type=int, default=8)
    ap.add_argument("--d", type=int, default=2)
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--oracle-gates", type=int, default=8)
    a = ap.parse_args()
    pkg = entry.import_package()
    from itnn_b200 import graphs, problems

    dtype = np.float64 if a.dtype == "f64" else np.complex128
    p = problems.synthetic_peps(graphs.named_grid(tuple(a.lattice)), a.chi, a.d, dtype)
    ga = p.ga
    rng = np.random.default_rng(0)
    # the four matchings of the square lattice (horizontal even/odd, vertical even/odd) = one Trotter step
    nx, ny = a.lattice
    vid = {v: i for i, v in enumerate(ga.vertices)}
    layers = []
    for axis in (0, 1):
        for parity in (0, 1):
            es = []
            for x in range(1, nx + 1):
                for y in range(1, ny + 1):
                    w = (x + 1, y) if axis == 0 else (x, y + 1)
                    if (x if axis == 0 else y) % 2 == parity and w in vid:
                        es.append(ga.edge_index[(vid[(x, y)], vid[w])])
            layers.append(es)
    dd = a.d * a.d

    def rand_ops(n):
        o = rng.standard_normal((n, dd * dd))
        if dtype == np.complex128:
            o = o + 1j * rng.standard_normal((n, dd * dd))
        # near-identity gates keep the state well conditioned over many layers
        return [(np.eye(dd).ravel() + 0.1 * o[i]).reshape((a.d,) * 4, order="F").astype(dtype) for i in range(n)]

    with pkg.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(3, 0.0, True)
        msgs0 = ctx.get_messages()
        times, gates = [], 0
        for it in range(a.warmup + a.layers):
            es = layers[it % 4]
            ops = rand_ops(len(es))
            t0 = time.perf_counter()
            ctx.apply_two_site_gates(es, ops, max_rank=a.chi, normalize=True)
            dt = time.perf_counter() - t0
            if it >= a.warmup:
                times.append(dt)
                gates += len(es)
        res, _ = ctx.sweep(1, 0.0, True)
        stats = ctx.apply_stats()
    out = {
        "metric": "bp_simple_update_gates_per_s", "value": gates / sum(times), "unit": "gates/s",
        "config": {"workload": f"{nx}x{ny} square-lattice PEPS, chi={a.chi}, d={a.d}, {a.dtype}; layers = the four matchings",
                   "gates_per_layer": [len(l) for l in layers], "layers_timed": a.layers,
                   "timing": "host wall clock around the synchronous C-ABI call"},
        "ms_per_layer": 1e3 * sum(times) / len(times), "residual_of_next_sweep": res,
        "gates_on_gram_kernel": stats[0], "gates_declined_to_stepwise_kernel": stats[1],
    }
    if a.oracle_gates > 0:  # CPU baseline: the numpy oracle of the reference algorithm, one thread, a few gates of layer 0
        from oracle import apply_oracle as A

        link = lambda v, w: ("l", min(v, w), max(v, w))
        state = {}
        for v in range(ga.nv):
            nb = [ga.dst[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
            state[v] = (np.asarray(p.tensors[v]), (("s", v),) + tuple(link(v, w) for w in nb))
        env = {(ga.src[e], ga.dst[e]): msgs0[e] for e in range(ga.ne)}
        es = layers[0][len(layers[0]) // 2:][: a.oracle_gates]  # bulk gates
        ops = rand_ops(len(es))
        t0 = time.perf_counter()
        for e, op in zip(es, ops):
            names = (("s", ga.src[e]), ("s", ga.dst[e]))
            A.apply_operator((op, names, names), state, env, trunc=a.chi, normalize=True)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": len(es) / dt, "unit": "gates/s", "cores": 1, "kind": "port",
                               "sample": f"{len(es)} bulk gates of layer 0, numpy oracle (oracle/apply_oracle.py)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
