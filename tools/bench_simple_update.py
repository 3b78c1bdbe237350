"""One imaginary-time simple-update step on a square-lattice PEPS, end to end on one context (for round 2's measurements):
four layers of vertex-disjoint two-site gates (bpx_apply_two_site_gates), `--sweeps` synchronous BP sweeps
(bpx_sweep_async), and the bond energies (bpx_edge_expect).  Prints ONE JSON line with the wall time of each phase.

  python tools/bench_simple_update.py [--lattice 32 32] [--chi 8] [--steps 5] [--sweeps 2] [--dtype f64|c128]

Transverse-field Ising gates exp(-dt h_e) with the field shared between the bonds of a vertex (tests/test_zzz_resident_state.py
checks the same loop against exact ground states on trees).  Not part of bench.py's contract."""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy.linalg import expm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def square_lattice_layers(ga, nx, ny):
    """The four matchings of the open square lattice (directed edge ids), horizontal even/odd then vertical even/odd."""
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
    return layers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=2, default=[32, 32])
    ap.add_argument("--chi", type=int, default=8)
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--field", type=float, default=3.0)
    a = ap.parse_args()
    pkg = entry.import_package()
    from itnn_b200 import graphs, problems

    dtype = np.float64 if a.dtype == "f64" else np.complex128
    nx, ny = a.lattice
    p = problems.synthetic_peps(graphs.named_grid((nx, ny)), a.chi, 2, dtype)
    ga = p.ga
    deg = np.diff(np.asarray(ga.row_ptr))
    X, Z, I2 = np.array([[0.0, 1.0], [1.0, 0.0]]), np.diag([1.0, -1.0]), np.eye(2)
    layers = square_lattice_layers(ga, nx, ny)

    def bond_h(e):
        return (-np.kron(Z, Z) - (a.field / deg[ga.src[e]]) * np.kron(X, I2) - (a.field / deg[ga.dst[e]]) * np.kron(I2, X))

    cache = {}

    def op_for(e, gate):
        key = (int(deg[ga.src[e]]), int(deg[ga.dst[e]]), gate)
        if key not in cache:
            h = bond_h(e)
            cache[key] = (expm(-a.dt * h) if gate else h).reshape(2, 2, 2, 2).astype(dtype)
        return cache[key]

    gate_ops = [[op_for(e, True) for e in es] for es in layers]
    all_edges = [e for es in layers for e in es]
    energy_ops = [op_for(e, False) for e in all_edges]
    t_gate, t_bp, t_energy, energies = [], [], [], []
    with pkg.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(5, 0.0, True)
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            for es, ops in zip(layers, gate_ops):
                ctx.apply_two_site_gates(es, ops, max_rank=a.chi, normalize=True)
            t1 = time.perf_counter()
            ctx.sweep_async(a.sweeps, True)
            res = ctx.last_residual()
            t2 = time.perf_counter()
            num, den = ctx.edge_expect(all_edges, energy_ops)
            t3 = time.perf_counter()
            if it >= a.warmup:
                t_gate.append(t1 - t0)
                t_bp.append(t2 - t1)
                t_energy.append(t3 - t2)
            energies.append(float(np.sum((num / den).real)) / ga.nv)
    n_gates = len(all_edges)
    step = np.mean(t_gate) + np.mean(t_bp) + np.mean(t_energy)
    print(json.dumps({
        "metric": "simple_update_steps_per_s", "value": 1.0 / step, "unit": "steps/s",
        "config": {"workload": f"{nx}x{ny} square-lattice PEPS, chi={a.chi}, d=2, {a.dtype}; TFI field {a.field}, dt {a.dt}",
                   "gates_per_step": n_gates, "bp_sweeps_per_step": a.sweeps, "updates_per_sweep": ga.ne, "steps_timed": a.steps,
                   "timing": "host wall clock per phase (every phase ends with a device synchronisation)"},
        "ms_per_step": 1e3 * step, "ms_gate_layers": 1e3 * float(np.mean(t_gate)), "ms_bp": 1e3 * float(np.mean(t_bp)),
        "ms_bond_energies": 1e3 * float(np.mean(t_energy)), "gates_per_s": n_gates / float(np.mean(t_gate)),
        "energy_per_site": energies, "residual_after_last_bp": res,
    }))


if __name__ == "__main__":
    main()
