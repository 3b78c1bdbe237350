#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

PROVENANCE.  The reference (pure Julia) holds no golden vectors for the BP path and cannot run in this image
(SURVEY.md §8 c1), so these fixtures are produced by the CPU ORACLE (oracle/bp_oracle.py), whose algorithm is pinned by
the reference's own known-answer tests (tests/test_oracle_known_answers.py, tests/test_apply_oracle.py,
tests/test_generators.py).  They are regression anchors: the numpy oracle, the C oracle and the CUDA path must all keep
reproducing them (tests/test_zz_golden.py).  Inputs come from the library's deterministic host RNG (bpx_fill_randn) or
from closed-form tensors, so nothing but this script is needed to rebuild them:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402

entry.import_package()
o = entry.import_oracle()
from helpers import spin_ice_tensors  # noqa: E402
from itnn_b200 import graphs, problems  # noqa: E402


def stack(msgs):
    return np.concatenate([np.asarray(m).ravel(order="F") for m in msgs])


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{path}: {os.path.getsize(path)} bytes")


def jacobi_case(name, p, nsweeps):
    """Per-sweep messages, residuals and beliefs of synchronous sweeps on a synthetic PEPS norm network."""
    op = o.make_problem(p.ga, p.tensors, "norm")
    msgs, per_sweep, res = list(p.messages), [], []
    for _ in range(nsweeps):
        prev, msgs = msgs, o.sweep_jacobi(op, msgs)
        per_sweep.append(stack(msgs))
        res.append(o.iterate_diff(msgs, prev))
    sz = np.diag([1.0, -1.0]).astype(p.dtype)
    save(name, src=np.asarray(p.ga.src), dst=np.asarray(p.ga.dst), slot=np.asarray(p.ga.slot),
         sites=stack(p.tensors), messages0=stack(p.messages), messages=np.stack(per_sweep), residual=np.array(res),
         vertex_scalars=np.array(o.vertex_scalars(op, msgs)), edge_scalars=np.array(o.edge_scalars(op, msgs)),
         expect_sz=np.array([o.local_expect(op, msgs, v, sz) for v in range(p.ga.nv)]),
         chi=p.chi, d=p.d)


def main():
    # BASELINE config 1: 4x4 open square lattice, chi = 2, d = 2, Float64 (the reference-runnable case)
    jacobi_case("cfg1_jacobi", problems.make_config("cfg1"), 4)
    # ComplexF64, mixed degrees 1..3: a 3 x 2 comb tree with chi = 3, d = 2
    g = graphs.named_comb_tree((3, 2))
    jacobi_case("comb32_c128_jacobi", problems.synthetic_peps(g, 3, 2, np.complex128, seed=7, name="comb"), 3)
    # reference schedule: spin ice on the 3x3 torus, sequential sweeps to tol 1e-10 (test/test_beliefpropagation.jl:204-225)
    g = graphs.named_grid((3, 3), periodic=True)
    ga = graphs.graph_arrays(g)
    p = o.make_problem(ga, spin_ice_tensors(ga), "single")
    from itnn_b200.device import fill_randn
    m0 = [np.abs(fill_randn(123, e, np.float64, 2)) % 1.0 for e in range(ga.ne)]
    seq = [ga.edge_id(e) for e in graphs.forest_cover_edge_sequence(g)]
    out, it, delta = o.beliefpropagation(p, m0, maxiter=10, tol=1e-10, schedule="sequential", edge_seq=seq)
    save("spin_ice_3x3_sequential", messages0=stack(m0), edge_seq=np.array(seq), messages=stack(out), iterations=it,
         delta=delta, log_z_bp=o.bethe_free_energy(p, out), log_z_exact=9 * np.log(1.5))
    # single layer, synchronous: the Ising generator's network on the 4x4 torus (bench.py --workload ising recipe)
    q = problems.synthetic_ising((4, 4), beta=0.3)
    tensors, msgs = problems.unpacked(q)
    p = o.make_problem(q.ga, tensors, "single")
    per_sweep, res = [], []
    for _ in range(3):
        prev, msgs = msgs, o.sweep_jacobi(p, msgs)
        per_sweep.append(stack(msgs))
        res.append(o.iterate_diff(msgs, prev))
    save("ising_4x4_torus_jacobi", sites=q.tensors, messages0=q.messages, messages=np.stack(per_sweep), residual=np.array(res),
         log_z_exact=np.log(o.contract_all_sequential(p)))


if __name__ == "__main__":
    main()
