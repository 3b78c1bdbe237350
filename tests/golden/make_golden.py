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

    apply_case()


def apply_case():
    """BP simple-update gate layer (src/apply/apply_operators.jl:246-283): a 3x3 ComplexF64 PEPS (chi = 3, d = 2) with the
    messages of four synchronous BP sweeps as environment, one layer of three disjoint two-site gates truncated to rank 2
    with S normalised, then one one-site gate.  Stored: inputs, kept singular values, and per gate the GAUGE-INVARIANT
    content of the result (the two new tensors contracted over their bond, axes sorted by name) -- the tensors
    themselves are only defined up to the phases of the singular vectors."""
    from oracle import apply_oracle as A

    p = problems.synthetic_peps(graphs.named_grid((3, 3)), 3, 2, np.complex128, seed=11, name="apply33")
    ga = p.ga
    op = o.make_problem(ga, p.tensors, "norm")
    msgs = list(p.messages)
    for _ in range(4):
        msgs = o.sweep_jacobi(op, msgs)
    state, env = apply_state(ga, p.tensors, msgs)
    from itnn_b200.device import fill_randn
    edges = [ga.edge_index[(0, 1)], ga.edge_index[(5, 4)], ga.edge_index[(6, 7)]]
    ops = [fill_randn(11, 1000 + i, np.complex128, 16).reshape((2, 2, 2, 2), order="F") for i in range(len(edges))]
    svs, pairs = [], []
    for e, g_ in zip(edges, ops):
        v1, v2 = ga.src[e], ga.dst[e]
        names = (("s", v1), ("s", v2))
        new_state, new_env = A.apply_operator((g_, names, names), state, env, trunc=2, normalize=True)
        svs.append(np.diag(new_env[(v1, v2)]).real)
        t = A.contract(new_state[v1], new_state[v2])
        pairs.append(A.permute(t, sorted(t[1], key=repr)).ravel(order="F"))
    one = fill_randn(11, 2000, np.complex128, 4).reshape((2, 2), order="F")
    names = (("s", 8),)
    one_state, _ = A.apply_operator((one, names, names), state, env, normalize=True)
    save("apply_grid33_c128", sites=stack(p.tensors), messages=stack(msgs), edges=np.array(edges), ops=np.stack(ops),
         max_rank=2, normalize=1, singular_values=np.stack(svs), pair_products=np.concatenate(pairs),
         pair_sizes=np.array([len(x) for x in pairs]), one_site_vertex=8, one_site_op=one,
         one_site_result=A.permute(one_state[8], state[8][1]).ravel(order="F"), chi=3, d=2, seed=11)


def apply_state(ga, tensors, msgs):
    """Canonical arrays -> the apply oracle's named state / environment (site name ("s", v), link name ("l", lo, hi))."""
    link = lambda v, w: ("l", min(v, w), max(v, w))
    state = {}
    for v in range(ga.nv):
        nb = [ga.dst[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
        state[v] = (np.asarray(tensors[v]), (("s", v),) + tuple(link(v, w) for w in nb))
    env = {(ga.src[e], ga.dst[e]): np.asarray(msgs[e]) for e in range(ga.ne)}
    return state, env


if __name__ == "__main__":
    main()
