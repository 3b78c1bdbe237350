"""ITensorNetworkGenerators mirror (itensornetworksnext.jl_b200/generators.py) against the reference's own
tests for it (test/test_itensornetworkgenerators.jl) and the analytic Ising free energies of test/utils.jl.
The exact contractions and the CPU BP runs use the oracle; the generators themselves are host bookkeeping.
SURVEY.md §8 c4 (7)."""
import math

import numpy as np
import pytest
from scipy.integrate import quad

from itnn_b200 import generators as gen
from itnn_b200 import graphs
from itnn_b200.tensornetwork import Index, canonical_arrays


# ---- test/utils.jl:4-39 -------------------------------------------------------------------------------
def beta_c_2d_ising():
    return math.log(1 + math.sqrt(2)) / 2


def f_2d_ising(beta, J=1.0):
    kappa = 2 * math.sinh(2 * beta * J) / math.cosh(2 * beta * J) ** 2
    integral, _ = quad(lambda th: math.log((1 + math.sqrt(abs(1 - (kappa * math.sin(th)) ** 2))) / 2), 0, math.pi)
    return (-math.log(2 * math.cosh(2 * beta * J)) - integral / (2 * math.pi)) / beta


def f_1d_ising_periodic(beta, n, J=1.0, h=0.0):
    r = math.sqrt(math.sinh(beta * h) ** 2 + math.exp(-4 * beta * J))
    lp = math.exp(beta * J) * (math.cosh(beta * h) + r)
    lm = math.exp(beta * J) * (math.cosh(beta * h) - r)
    return -math.log(lp ** n + lm ** n) / (beta * n)


def f_1d_ising_open(beta, n, J=1.0, h=0.0):
    t = np.array([[math.exp(beta * (J + h)), math.exp(-beta * J)], [math.exp(-beta * J), math.exp(beta * (J - h))]])
    b = np.array([math.exp(beta * h / 2), math.exp(-beta * h / 2)])
    return -math.log(b @ np.linalg.matrix_power(t, n - 1) @ b) / (beta * n)


def link_function(g, dim=2):
    ldict = {frozenset((e.src, e.dst)): Index(dim) for e in g.edges()}
    return lambda e: ldict[frozenset((e.src, e.dst))]


def contract(oracle, tn):
    cp = canonical_arrays(tn)
    p = oracle.make_problem(cp.ga, cp.tensors, "single")
    z = oracle.contract_all_sequential(p)
    if cp.ga.nv <= 9:  # the two exact contractions of the oracle agree (the einsum one is slow on larger loopy graphs)
        assert np.isclose(z, oracle.contract_all(p), rtol=1e-12)
    return z


# ---- "Delta Network": test/test_itensornetworkgenerators.jl:14-28 ---------------------------------------
def test_delta_network():
    g = graphs.named_grid((3, 3))
    l = link_function(g)
    tn = gen.delta_network(l, g)
    assert len(tn.vertices()) == 9
    assert len(tn.edges()) == g.ne()
    assert set(tn.vertices()) == set(g.vertices())
    assert {frozenset((e.src, e.dst)) for e in tn.edges()} == {frozenset((e.src, e.dst)) for e in g.edges()}
    for v in tn.vertices():
        inds = [l(e) for e in g.incident_edges(v)]
        want = gen.delta(np.float64, inds)
        assert tn[v].inds == want.inds and np.array_equal(tn[v].data, want.data)
        assert tn[v].data.sum() == 2.0 and tn[v].data[(0,) * len(inds)] == 1.0 and tn[v].data[(1,) * len(inds)] == 1.0


def test_diagonaltensor_rectangular():
    t = gen.diagonaltensor([1.0, 2.0], [Index(2), Index(3), Index(2)])
    assert t.data[0, 0, 0] == 1.0 and t.data[1, 1, 1] == 2.0 and np.count_nonzero(t.data) == 2


def test_sqrt_ising_bond_squares_to_the_boltzmann_matrix():
    for beta, J, h, d1, d2 in [(0.4, 1.0, 0.0, 2, 2), (0.7, 0.5, 0.3, 3, 4), (0.2, 2.0, -0.4, 1, 2)]:
        m = gen.sqrt_ising_bond(beta, J, h, deg1=d1, deg2=d2)
        h1, h2 = h / d1, h / d2
        want = np.array([[math.exp(beta * (J + h1 + h2)), math.exp(beta * (-J + h1 - h2))],
                         [math.exp(beta * (-J - h1 + h2)), math.exp(beta * (J - h1 - h2))]])
        assert np.allclose(m @ m, want, rtol=1e-12)
    with pytest.raises(ValueError):
        gen.sqrt_ising_bond(0.4, -1.0, deg1=2, deg2=2)  # antiferromagnetic bond: DomainError in the reference


# ---- "1D Ising": test/test_itensornetworkgenerators.jl:30-54 --------------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_ising_1d_matches_analytic_free_energy(oracle, periodic):
    beta = 0.4
    g = graphs.named_grid((4,), periodic=periodic)
    l = link_function(g)
    tn = gen.ising_network(l, beta, g)
    assert len(tn.vertices()) == 4 and len(tn.edges()) == g.ne()
    for v in tn.vertices():
        inds = [l(e) for e in g.incident_edges(v)]
        assert set(tn[v].inds) == set(inds)
        assert not np.array_equal(tn[v].data, gen.delta(np.float64, inds).data)
    z = contract(oracle, tn)
    f = -math.log(z) / (beta * g.nv())
    want = f_1d_ising_periodic(beta, 4) if periodic else f_1d_ising_open(beta, 4)
    assert np.isclose(f, want)


# ---- "2D Ising": test/test_itensornetworkgenerators.jl:55-75 --------------------------------------------
def test_ising_2d_periodic_4x4_close_to_onsager(oracle):
    beta = beta_c_2d_ising()
    g = graphs.named_grid((4, 4), periodic=True)
    l = link_function(g)
    tn = gen.ising_network(l, beta, g)
    assert len(tn.vertices()) == 16 and len(tn.edges()) == g.ne() == 32
    z = contract(oracle, tn)
    f = -math.log(z) / (beta * g.nv())
    assert np.isclose(f, f_2d_ising(beta), rtol=1e-1)
    # brute-force partition function over the 2^16 spin configurations: the generator is exact, not just "close"
    verts = g.vertices()
    idx = {v: i for i, v in enumerate(verts)}
    spins = 1 - 2 * ((np.arange(1 << 16)[:, None] >> np.arange(16)) & 1)
    energy = sum(spins[:, idx[e.src]] * spins[:, idx[e.dst]] for e in g.edges())
    assert np.isclose(z, np.exp(beta * energy).sum(), rtol=1e-12)


# ---- field and sigma^z insertions (ising_network.jl:27-40) ---------------------------------------------
def test_ising_field_and_sz_vertices_against_brute_force(oracle):
    # (uniform degree: with a field the reference's bond matrix is symmetric only when both ends have the same degree)
    beta, J, h = 0.3, 0.8, 0.25
    g = graphs.named_grid((5,), periodic=True)
    l = link_function(g)
    verts = g.vertices()
    idx = {v: i for i, v in enumerate(verts)}
    n = len(verts)
    spins = 1 - 2 * ((np.arange(1 << n)[:, None] >> np.arange(n)) & 1)
    energy = J * sum(spins[:, idx[e.src]] * spins[:, idx[e.dst]] for e in g.edges()) + h * spins.sum(axis=1)
    w = np.exp(beta * energy)
    z = contract(oracle, gen.ising_network(l, beta, g, J=J, h=h))
    assert np.isclose(z, w.sum(), rtol=1e-12)
    assert np.isclose(-math.log(z) / (beta * n), f_1d_ising_periodic(beta, n, J=J, h=h), rtol=1e-12)
    v0 = verts[2]
    zs = contract(oracle, gen.ising_network(l, beta, g, J=J, h=h, sz_vertices=[v0]))
    assert np.isclose(zs / z, (w * spins[:, idx[v0]]).sum() / w.sum(), rtol=1e-12)


# ---- BP on the generators' networks: exact on trees (open chain), single-layer mode of the path ---------
@pytest.mark.parametrize("n,h", [(4, 0.0), (9, 0.0)])
def test_bp_on_open_ising_chain_gives_the_analytic_free_energy(oracle, n, h):
    beta = 0.4
    g = graphs.named_grid((n,))
    tn = gen.ising_network(link_function(g), beta, g, h=h)
    cp = canonical_arrays(tn)
    p = oracle.make_problem(cp.ga, cp.tensors, "single")
    msgs = [np.ones(2) for _ in range(cp.ga.ne)]
    for _ in range(n):  # Jacobi sweeps: exact after diameter-many
        msgs = oracle.sweep_jacobi(p, msgs)
    f_bp = -oracle.bethe_free_energy(p, msgs) / (beta * n)
    assert np.isclose(f_bp, f_1d_ising_open(beta, n, h=h), rtol=1e-12)
