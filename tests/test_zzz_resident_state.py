"""`ResidentState` (itensornetworksnext.jl_b200/resident.py): gates, BP and expectation values on ONE device context.

Known answer, end to end: imaginary-time evolution of the transverse-field Ising model H = -sum ZZ - h sum X on an open
chain / a comb tree by BP simple update (two-site gates exp(-dt h_bond), truncated to the fixed chi, S normalised).  On a
TREE the BP environments are exact, chi is large enough to hold the exact state, so the evolved state must reach the
exact ground state (energy up to the Trotter error, computed here by brute force from the downloaded tensors), and the
BP local expectation values of the resident state must equal the exact ones of that state to 1e-9 (BASELINE.json's
tolerance for converged local expectation values).

  * `-m "not gpu"`: the driver logic with tests/native_ctx.HostHarnessContext (gate kernels compiled for the host, BP from
    the numpy oracle) -- test infrastructure only.
  * `-m gpu`: everything on the B200 through libbpx.so.
"""
import numpy as np
import pytest
from scipy.linalg import expm

import itnn_b200 as B
from itnn_b200 import apply as apply_mod
from itnn_b200 import graphs
from native_ctx import HostHarnessContext, build_hostlib
from oracle import apply_oracle as A

X = np.array([[0.0, 1.0], [1.0, 0.0]])
Z = np.diag([1.0, -1.0])
I2 = np.eye(2)


def product_state(g, chi, dtype=np.float64):
    """|+>^n with every link padded to `chi` (only bond index 0 is populated)."""
    link = {frozenset((e.src, e.dst)): B.Index(chi, ("l", e.src, e.dst)) for e in g.edges()}
    tensors = {}
    for v in g.vertices():
        inds = [B.Index(2, ("s", v))] + [link[frozenset((v, w))] for w in g.neighbors(v)]
        data = np.zeros([i.dim for i in inds], dtype=dtype)
        data[(slice(None),) + (0,) * (len(inds) - 1)] = np.array([1.0, 1.0]) / np.sqrt(2.0)
        tensors[v] = B.ITensor(data, inds)
    return B.ITensorNetwork(tensors)


def bond_gates(g, h, dt, dtype=np.float64):
    """exp(-dt h_e), h_e = -Z Z - (h / z_v) X 1 - (h / z_w) 1 X: the field is shared by the bonds of a vertex."""
    ops = []
    for e in g.edges():
        zv, zw = g.degree(e.src), g.degree(e.dst)
        hb = -np.kron(Z, Z) - (h / zv) * np.kron(X, I2) - (h / zw) * np.kron(I2, X)
        u = expm(-dt * hb).reshape(2, 2, 2, 2)  # C order: [o_src, o_dst, i_src, i_dst]
        names = (("s", e.src), ("s", e.dst))
        ops.append((e, B.Operator(u.astype(dtype), names, names)))
    # order the gates layer by layer (greedy edge colouring): consecutive disjoint gates then share one device call
    layers = []
    for e, op in ops:
        for layer in layers:
            if all(not ({e.src, e.dst} & {f.src, f.dst}) for f, _ in layer):
                layer.append((e, op))
                break
        else:
            layers.append([(e, op)])
    return [op for layer in layers for _, op in layer]


def exact_ground_state(g, h):
    vs = g.vertices()
    n = len(vs)
    idx = {v: i for i, v in enumerate(vs)}

    def site_op(o, i):
        out = np.ones((1, 1))
        for j in range(n):
            out = np.kron(out, o if j == i else I2)
        return out

    ham = np.zeros((2 ** n, 2 ** n))
    for e in g.edges():
        ham -= site_op(Z, idx[e.src]) @ site_op(Z, idx[e.dst])
    for v in vs:
        ham -= h * site_op(X, idx[v])
    w, u = np.linalg.eigh(ham)
    return ham, w[0], u[:, 0]


def dense_vector(net, g):
    state = {v: (net[v].data, net[v].dimnames()) for v in net.vertices()}
    full = A.permute(A.prod(state), [("s", v) for v in g.vertices()])
    return full.reshape(-1)  # C order: first vertex slowest, like np.kron in exact_ground_state


def evolve_and_check(g, chi, h=1.0, steps=(60, 60, 60), dts=(0.1, 0.03, 0.01)):
    ham, e0, psi0 = exact_ground_state(g, h)
    with B.ResidentState(product_state(g, chi)) as rs:
        info = rs.beliefpropagation(dict(maxiter=50, tol=1e-14), schedule="sequential")
        assert info.delta < 1e-12
        for n, dt in zip(steps, dts):
            gates = bond_gates(g, h, dt)
            for _ in range(n):
                rs.apply_operators(gates, trunc=chi, normalize=True)  # layers of disjoint gates share a device call
        info = rs.beliefpropagation(dict(maxiter=50, tol=1e-14), schedule="sequential")
        assert info.delta < 1e-12
        ex = np.array(rs.expect(X)).real
        ez = np.array(rs.expect(Z)).real
        net = rs.state()
    psi = dense_vector(net, g)
    psi = psi / np.linalg.norm(psi)
    energy = psi @ ham @ psi
    assert e0 - 1e-10 <= energy <= e0 + 2e-3 * abs(e0), (energy, e0)         # Trotter error of dt = 0.01
    assert abs(abs(psi @ psi0) - 1.0) < 1e-3
    # BP beliefs on a tree are exact: local expectation values of the resident state == brute force on ITS tensors
    n = len(g.vertices())
    for i in range(n):
        def site_op(o):
            out = np.ones((1, 1))
            for j in range(n):
                out = np.kron(out, o if j == i else I2)
            return out

        assert abs(ex[i] - psi @ site_op(X) @ psi) < 1e-9
        assert abs(ez[i] - psi @ site_op(Z) @ psi) < 1e-9
    assert np.all(ex > 0.3)  # h = 1: strongly polarised along X, no symmetry breaking


def bond_hamiltonians(g, h):
    """h_e = -Z Z - (h / z_v) X 1 - (h / z_w) 1 X as two-site operators: sum_e h_e = H."""
    ops = []
    for e in g.edges():
        zv, zw = g.degree(e.src), g.degree(e.dst)
        hb = -np.kron(Z, Z) - (h / zv) * np.kron(X, I2) - (h / zw) * np.kron(I2, X)
        names = (("s", e.src), ("s", e.dst))
        ops.append(B.Operator(hb.reshape(2, 2, 2, 2), names, names))
    return ops


def check_energy_from_two_site_expectations(g, chi=4, h=1.0):
    """sum_e <h_e> from the device (bpx_edge_expect in the BP environment) == <H> by brute force on the downloaded
    tensors: exact on a tree.  A short evolution makes the state entangled and the environments non-trivial."""
    ham, e0, _ = exact_ground_state(g, h)
    with B.ResidentState(product_state(g, chi)) as rs:
        rs.beliefpropagation(dict(maxiter=50, tol=1e-14), schedule="sequential")
        gates = bond_gates(g, h, 0.1)
        for _ in range(15):
            rs.apply_operators(gates, trunc=chi, normalize=True)
        info = rs.beliefpropagation(dict(maxiter=50, tol=1e-14), schedule="sequential")
        assert info.delta < 1e-12
        bond_energies = np.array(rs.expect_two_site(bond_hamiltonians(g, h)))
        net = rs.state()
    psi = dense_vector(net, g)
    psi = psi / np.linalg.norm(psi)
    energy = psi @ ham @ psi
    assert np.abs(bond_energies.imag).max() < 1e-12 if np.iscomplexobj(bond_energies) else True
    assert abs(bond_energies.real.sum() - energy) < 1e-9 * abs(energy)
    assert e0 - 1e-10 <= energy < 0.9 * e0  # variational, and already close to the ground state after 15 steps


@pytest.fixture
def host_ctx(monkeypatch):
    lib = build_hostlib()
    made = []

    def factory(device=0):
        c = HostHarnessContext(lib, device)
        made.append(c)
        return c

    monkeypatch.setattr(apply_mod, "BPXContext", factory)
    return made


def test_tfi_chain_ground_state_host_harness(host_ctx):
    evolve_and_check(graphs.named_path_graph(4), chi=4, steps=(30, 30, 40))
    (ctx,) = host_ctx  # ONE context for the whole evolution
    # every Trotter step = two device calls: bonds (1,2),(3,4) share a layer, bond (2,3) is the next
    assert ctx.calls[:4] == [("two", 2), ("two", 1), ("two", 2), ("two", 1)]


def test_resident_state_argument_errors(host_ctx):
    g = graphs.named_path_graph(3)
    with B.ResidentState(product_state(g, 1)) as rs:
        gate = bond_gates(g, 1.0, 0.1)[0]
        with pytest.raises(B.ArgumentError, match="grow its bond"):
            rs.apply_layer([gate])  # chi = 1 cannot hold the gated bond and a resident state keeps its dimensions
        with pytest.raises(B.ArgumentError, match="vertex-disjoint"):
            rs.apply_layer(bond_gates(g, 1.0, 0.1), trunc=1)
        one = B.Operator(X, [("s", 1)], [("s", 1)])
        with pytest.raises(B.ArgumentError, match="not both"):
            rs.apply_layer([one, bond_gates(g, 1.0, 0.1)[1]], trunc=1)
        assert rs.apply_layer([]) == []
        rs.apply_layer([one])
        assert np.allclose(rs.state()[1].data.ravel(), np.array([1.0, 1.0]) / np.sqrt(2))  # X |+> = |+>
        env = rs.env()
        assert len(env) == 4 and all(m.data.shape == (1, 1) for m in env.values())


@pytest.mark.gpu
def test_tfi_chain_ground_state_gpu():
    evolve_and_check(graphs.named_path_graph(4), chi=4, steps=(30, 30, 40))


@pytest.mark.gpu
def test_tfi_comb_tree_ground_state_gpu():
    """Degree-3 vertices: the 3 x 2 comb tree (6 sites), chi = 4."""
    evolve_and_check(graphs.named_comb_tree((3, 2)), chi=4, steps=(30, 30, 40))


def test_tfi_comb_tree_ground_state_host_harness(host_ctx):
    evolve_and_check(graphs.named_comb_tree((3, 2)), chi=4, steps=(30, 30, 40))


@pytest.mark.parametrize("graph", ["chain", "comb"])
def test_energy_from_two_site_expectations_host_harness(host_ctx, graph):
    check_energy_from_two_site_expectations(graphs.named_path_graph(5) if graph == "chain" else graphs.named_comb_tree((3, 2)))
