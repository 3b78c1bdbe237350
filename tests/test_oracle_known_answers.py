"""Pins the CPU oracle (oracle/bp_oracle.py, oracle/bp_oracle.c) against every known-answer test the
reference's own suite holds for the BP path.  No GPU needed."""
import numpy as np
import pytest

from helpers import peps_tensors, positive_messages, randn, rel_err, single_layer_tensors, spin_ice_tensors
from itnn_b200 import graphs
from oracle.c_oracle import COracle

DTYPES = [np.float64, np.complex128]


def _seq(g, ga):
    return [ga.edge_id(e) for e in graphs.forest_cover_edge_sequence(g)]


# -- (1) tree exactness in ONE sequential sweep: test/test_beliefpropagation.jl:157-202 ---------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("gname,chi", [("chain2", 2), ("comb43", 3)])
def test_tree_exact_single_layer(oracle, dtype, gname, chi):
    g = graphs.named_grid((2, 1)) if gname == "chain2" else graphs.named_comb_tree((4, 3))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(123)
    p = oracle.make_problem(ga, single_layer_tensors(ga, chi, dtype, rng), "single")
    msgs = [np.ones(chi, dtype=dtype) for _ in range(ga.ne)]
    out, it, _ = oracle.beliefpropagation(p, msgs, maxiter=1, schedule="sequential", edge_seq=_seq(g, ga))
    assert it == 1
    z_bp = np.exp(oracle.bethe_free_energy(p, out))
    z_exact = oracle.contract_all(p)
    assert np.isclose(z_bp, z_exact, rtol=np.finfo(np.float64).eps ** (1 / 3))


def test_forest_cover_sequence_covers_every_directed_edge_once():
    for g in (graphs.named_grid((4, 4)), graphs.named_grid((3, 3), periodic=True), graphs.heavy_hex_127(),
              graphs.named_comb_tree((4, 3))):
        seq = graphs.forest_cover_edge_sequence(g)
        assert len(seq) == 2 * g.ne()
        assert len(set(seq)) == len(seq)
        assert set(seq) == set(g.all_edges())


# -- (2) spin ice: z_bp = 1.5^(n^2): test/test_beliefpropagation.jl:204-225 -------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 4, 5])
def test_spin_ice(oracle, dtype, n, schedule="sequential"):
    g = graphs.named_grid((n, n), periodic=True)
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(123)
    p = oracle.make_problem(ga, [t.astype(dtype) for t in spin_ice_tensors(ga)], "single")
    msgs = [rng.random(2).astype(dtype) if np.dtype(dtype).kind != "c" else
            (rng.random(2) + 1j * rng.random(2)).astype(dtype) for _ in range(ga.ne)]
    # Reference schedule (sequential).  NOTE: the synchronous (Jacobi) schedule does not converge on this
    # model (the uniform fixed point is marginal and it oscillates at residual ~4e-3), so spin ice pins
    # the sequential schedule only; the Jacobi schedule is pinned on trees below.
    out, it, delta = oracle.beliefpropagation(p, msgs, maxiter=10, tol=1e-10, schedule=schedule,
                                             edge_seq=_seq(g, ga))
    z_bp = np.exp(oracle.bethe_free_energy(p, out))
    assert np.isclose(z_bp, 1.5 ** (n * n))


# -- (3) iterate_diff(c, copy(c)) ~ 0: test/test_beliefpropagation.jl:134-150 -----------------------
def test_iterate_diff_identical(oracle):
    g = graphs.named_grid((2,))
    ga = graphs.graph_arrays(g)
    msgs = [np.ones(2) for _ in range(ga.ne)]
    assert abs(oracle.iterate_diff(msgs, [m.copy() for m in msgs])) <= 10 * np.finfo(float).eps


# -- (4) incoming messages exclude the reverse edge: test/test_beliefpropagation.jl:104-114 ---------
def test_incoming_exclusion_rule(oracle):
    g = graphs.named_path_graph(3)
    ga = graphs.graph_arrays(g)
    p = oracle.make_problem(ga, single_layer_tensors(ga, 2, np.float64, np.random.default_rng(0)), "single")
    e23, e12, e21, e32 = (ga.edge_id(x) for x in ((2, 3), (1, 2), (2, 1), (3, 2)))
    assert [f for f in oracle.incoming_edges(p, e23) if f is not None] == [e12]
    assert [f for f in oracle.incoming_edges(p, e12) if f is not None] == []
    assert [f for f in oracle.incoming_edges(p, e21) if f is not None] == [e32]


# -- (5) NormNetwork: <psi|psi> = ||prod(tn)||^2, real, positive (test/test_normnetwork.jl:148-166);
#        with (1): BP on a TREE norm network gives prod(vertex scalars)/prod(edge scalars) = <psi|psi> --
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("gname", ["path4", "comb32", "star"])
def test_norm_network_tree_exact(oracle, dtype, gname):
    if gname == "path4":
        g = graphs.named_path_graph(4)
    elif gname == "comb32":
        g = graphs.named_comb_tree((3, 2))
    else:
        g = graphs.NamedGraph(range(5))
        for w in range(1, 5):
            g.add_edge(0, w)
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(123)
    p = oracle.make_problem(ga, peps_tensors(ga, 3, 2, dtype, rng), "norm")
    msgs = [np.ones((3, 3), dtype=dtype) for _ in range(ga.ne)]
    out, _, _ = oracle.beliefpropagation(p, msgs, maxiter=1, schedule="sequential", edge_seq=_seq(g, ga))
    z_bp = np.exp(oracle.bethe_free_energy(p, out))
    z_exact = oracle.contract_all(p)
    assert abs(z_exact.imag) < 1e-12 * abs(z_exact) and z_exact.real > 0
    assert np.isclose(z_bp, z_exact.real, rtol=1e-9)
    # messages of a norm network stay Hermitian (bra/ket symmetric) when started Hermitian
    for m in out:
        assert np.allclose(m, m.conj().T, atol=1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
def test_jacobi_tree_exact_after_diameter_sweeps(oracle, dtype):
    """On a tree the synchronous schedule is exact once information crossed the diameter."""
    g = graphs.named_comb_tree((4, 3))  # diameter 7
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(11)
    for mode, tensors, chi in (("single", single_layer_tensors(ga, 3, dtype, rng), 3),
                               ("norm", peps_tensors(ga, 2, 2, dtype, rng), 2)):
        p = oracle.make_problem(ga, tensors, mode)
        shape = (chi, chi) if mode == "norm" else (chi,)
        msgs = [np.ones(shape, dtype=dtype) for _ in range(ga.ne)]
        out, it, delta = oracle.beliefpropagation(p, msgs, maxiter=8, schedule="jacobi")
        z_bp = np.exp(oracle.bethe_free_energy(p, out))
        assert np.isclose(z_bp, oracle.contract_all(p), rtol=1e-9)
        again = oracle.sweep_jacobi(p, out)
        assert oracle.iterate_diff(again, out) < 1e-13


# -- (6) the literal double-layer evaluation equals the absorption order; numpy == C ------
@pytest.mark.parametrize("dtype", DTYPES)
def test_literal_equals_absorption_and_c_equals_numpy(oracle, dtype):
    g = graphs.named_grid((3, 3))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(7)
    link_dim = [0] * ga.ne
    for e in range(ga.ne):  # non-uniform link dims 2/3
        link_dim[e] = link_dim[ga.rev[e]] = 2 + (min(e, ga.rev[e]) % 2)
    tensors = peps_tensors(ga, None, 2, dtype, rng, link_dim)
    p = oracle.make_problem(ga, tensors, "norm")
    msgs = positive_messages(ga, link_dim, dtype, rng)
    a = oracle.sweep_jacobi(p, msgs)
    b = oracle.sweep_jacobi(p, msgs, literal=True)
    assert rel_err(a, b) < 1e-12
    co = COracle(ga, [2] * ga.nv, link_dim, tensors, dtype)
    c = co.unpack(co.sweep_jacobi(co.pack(msgs)))
    cl = co.unpack(co.sweep_jacobi(co.pack(msgs), variant=1))
    assert rel_err(c, a) < 1e-12 and rel_err(cl, a) < 1e-12
    seq = _seq(g, ga)
    s_np = oracle.sweep_sequential(p, msgs, seq)
    s_c = co.unpack(co.sweep_sequential(co.pack(msgs), seq))
    assert rel_err(s_c, s_np) < 1e-12
    assert abs(co.iterate_diff(co.pack(a), co.pack(msgs)) - oracle.iterate_diff(a, msgs)) < 1e-13


# -- (7) zero-sum guard and normalize = false (beliefpropagation.jl:248-253) --------------------------
def test_zero_sum_guard(oracle):
    m = np.array([[1.0, -1.0], [2.0, -2.0]])
    assert np.array_equal(oracle.normalize_message(m), m)
    m2 = np.array([[1.0, 1.0], [2.0, 0.0]])
    assert np.isclose(oracle.normalize_message(m2).sum(), 1.0)


# -- (8) synchronous and sequential schedules share their fixed point (SURVEY §7 hard part 1) ---------
@pytest.mark.parametrize("dtype", DTYPES)
def test_schedules_share_fixed_point(oracle, dtype):
    g = graphs.named_grid((3, 3))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(5)
    p = oracle.make_problem(ga, peps_tensors(ga, 2, 2, dtype, rng), "norm")
    msgs = positive_messages(ga, [2] * ga.ne, dtype, rng)
    a, ita, _ = oracle.beliefpropagation(p, msgs, maxiter=400, tol=1e-15, schedule="jacobi")
    b, itb, _ = oracle.beliefpropagation(p, msgs, maxiter=400, tol=1e-15, schedule="sequential", edge_seq=_seq(g, ga))
    sz = np.diag([1.0, -1.0]).astype(dtype)
    ea = [oracle.local_expect(p, a, v, sz) for v in range(ga.nv)]
    eb = [oracle.local_expect(p, b, v, sz) for v in range(ga.nv)]
    assert np.allclose(ea, eb, atol=1e-9)
