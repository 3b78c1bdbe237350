"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle -- the reference
has no golden vectors for this path and cannot run here): the numpy oracle, the C oracle and the CUDA path must all keep
reproducing them.  The fixtures carry their own inputs; the graphs are rebuilt from the committed edge arrays."""
import os

import numpy as np
import pytest

import itnn_b200 as B
from helpers import spin_ice_tensors
from itnn_b200 import _lib, graphs, problems
from oracle.c_oracle import COracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10  # BASELINE.json: per-sweep messages within 1e-10 relative


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def norm_problem(f, g):
    """(ga, tensors, messages0) of a norm-network fixture: the graph is rebuilt, the packed arrays are cut by shape."""
    ga = graphs.graph_arrays(g)
    assert np.array_equal(ga.src, f["src"]) and np.array_equal(ga.dst, f["dst"]) and np.array_equal(ga.slot, f["slot"])
    chi, d = int(f["chi"]), int(f["d"])
    deg = np.diff(ga.row_ptr)
    off = np.concatenate([[0], np.cumsum(d * chi ** deg)])
    tensors = [f["sites"][off[v]:off[v + 1]].reshape((d,) + (chi,) * int(deg[v]), order="F") for v in range(ga.nv)]
    msgs = [f["messages0"][chi * chi * e:chi * chi * (e + 1)].reshape((chi, chi), order="F") for e in range(ga.ne)]
    return ga, chi, d, tensors, msgs


NORM_CASES = {"cfg1_jacobi": lambda: graphs.named_grid((4, 4)), "comb32_c128_jacobi": lambda: graphs.named_comb_tree((3, 2))}


@pytest.mark.parametrize("name", sorted(NORM_CASES))
def test_oracles_reproduce_norm_network_fixtures(oracle, name):
    f = load(name)
    ga, chi, d, tensors, msgs = norm_problem(f, NORM_CASES[name]())
    p = oracle.make_problem(ga, tensors, "norm")
    co = COracle(ga, [d] * ga.nv, [chi] * ga.ne, tensors, tensors[0].dtype)
    flat = co.pack(msgs)
    for k in range(f["messages"].shape[0]):
        prev, msgs = msgs, oracle.sweep_jacobi(p, msgs)
        flat = co.sweep_jacobi(flat)
        got = np.concatenate([m.ravel(order="F") for m in msgs])
        assert rel(got, f["messages"][k]) < 1e-13
        assert rel(flat, f["messages"][k]) < 1e-12
        assert abs(oracle.iterate_diff(msgs, prev) - f["residual"][k]) < 1e-14
    assert np.allclose(oracle.vertex_scalars(p, msgs), f["vertex_scalars"], rtol=1e-13)
    assert np.allclose(oracle.edge_scalars(p, msgs), f["edge_scalars"], rtol=1e-13)
    if name == "cfg1_jacobi":  # the fixture's inputs ARE BASELINE config 1 from the shared RNG
        q = problems.make_config("cfg1")
        assert np.array_equal(np.concatenate([t.ravel(order="F") for t in q.tensors]), f["sites"])


def test_oracle_reproduces_single_layer_fixtures(oracle):
    f = load("spin_ice_3x3_sequential")
    g = graphs.named_grid((3, 3), periodic=True)
    ga = graphs.graph_arrays(g)
    p = oracle.make_problem(ga, spin_ice_tensors(ga), "single")
    m0 = [f["messages0"][2 * e:2 * e + 2] for e in range(ga.ne)]
    assert [ga.edge_id(e) for e in graphs.forest_cover_edge_sequence(g)] == list(f["edge_seq"])
    out, it, delta = oracle.beliefpropagation(p, m0, maxiter=10, tol=1e-10, schedule="sequential", edge_seq=list(f["edge_seq"]))
    assert it == int(f["iterations"]) and rel(np.concatenate(out), f["messages"]) < 1e-13
    assert np.isclose(float(f["log_z_bp"]), float(f["log_z_exact"]), rtol=1e-8)  # z_bp = 1.5^9
    f = load("ising_4x4_torus_jacobi")
    q = problems.synthetic_ising((4, 4), beta=0.3)
    assert np.allclose(q.tensors, f["sites"], rtol=1e-15) and np.array_equal(q.messages, f["messages0"])
    tensors, msgs = problems.unpacked(q)
    p = oracle.make_problem(q.ga, tensors, "single")
    for k in range(3):
        msgs = oracle.sweep_jacobi(p, msgs)
        assert rel(np.concatenate(msgs), f["messages"][k]) < 1e-13


# ---- the CUDA path against the same fixtures --------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [_lib.BPX_KERNEL_AUTO, _lib.BPX_KERNEL_GENERIC])
@pytest.mark.parametrize("name", sorted(NORM_CASES))
def test_gpu_reproduces_norm_network_fixtures(name, kernel):
    f = load(name)
    ga, chi, d, tensors, msgs = norm_problem(f, NORM_CASES[name]())
    sz = np.diag([1.0, -1.0]).astype(tensors[0].dtype)
    with B.BPXContext(0) as ctx:
        ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        ctx.set_kernel_policy(kernel)
        ctx.set_dims(tensors[0].dtype, "norm", [d] * ga.nv, [chi] * ga.ne)
        ctx.set_site_tensors(tensors)
        ctx.set_messages(msgs)
        for k in range(f["messages"].shape[0]):
            res, done = ctx.sweep(1)
            assert done == 1 and rel(ctx.get_messages_flat(), f["messages"][k]) < TOL
            assert abs(res - f["residual"][k]) < 1e-11
        assert np.allclose(ctx.vertex_scalars(), f["vertex_scalars"], rtol=1e-10)
        assert np.allclose(ctx.edge_scalars(), f["edge_scalars"], rtol=1e-10)
        num = ctx.vertex_expect_numerators([sz] * ga.nv)
        assert np.abs(num / ctx.vertex_scalars() - f["expect_sz"]).max() < 1e-9


@pytest.mark.gpu
def test_gpu_reproduces_single_layer_fixtures():
    f = load("spin_ice_3x3_sequential")
    g = graphs.named_grid((3, 3), periodic=True)
    ga = graphs.graph_arrays(g)
    with B.BPXContext(0) as ctx:
        ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        ctx.set_dims(np.float64, "single", None, [2] * ga.ne)
        ctx.set_site_tensors(spin_ice_tensors(ga))
        ctx.set_messages(f["messages0"].copy())
        res, done = ctx.sweep_sequence([int(e) for e in f["edge_seq"]], 10, 1e-10)
        assert done == int(f["iterations"]) and rel(ctx.get_messages_flat(), f["messages"]) < TOL
    f = load("ising_4x4_torus_jacobi")
    q = problems.synthetic_ising((4, 4), beta=0.3)
    for kernel in (_lib.BPX_KERNEL_AUTO, _lib.BPX_KERNEL_GENERIC):
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, q, kernel)
            for k in range(3):
                res, _ = ctx.sweep(1)
                assert rel(ctx.get_messages_flat(), f["messages"][k]) < TOL
                assert abs(res - f["residual"][k]) < 1e-11


# ---- gate application (src/apply/apply_operators.jl:246-283): fixture apply_grid33_c128 ----------------------
def _apply_fixture():
    from golden.make_golden import apply_state  # the committed generating script also defines the naming

    f = load("apply_grid33_c128")
    ga = graphs.graph_arrays(graphs.named_grid((3, 3)))
    chi, d = int(f["chi"]), int(f["d"])
    deg = np.diff(ga.row_ptr)
    off = np.concatenate([[0], np.cumsum(d * chi ** deg)])
    tensors = [f["sites"][off[v]:off[v + 1]].reshape((d,) + (chi,) * int(deg[v]), order="F") for v in range(ga.nv)]
    msgs = [f["messages"][chi * chi * e:chi * chi * (e + 1)].reshape((chi, chi), order="F") for e in range(ga.ne)]
    return f, ga, tensors, msgs, apply_state


def _check_apply_result(f, ga, apply_state, new_tensors, new_msgs, svs):
    from oracle import apply_oracle as A

    k = int(f["max_rank"])
    state, _ = apply_state(ga, new_tensors, new_msgs)
    pair_off = np.concatenate([[0], np.cumsum(f["pair_sizes"])])
    for i, e in enumerate(f["edges"]):
        v1, v2 = ga.src[e], ga.dst[e]
        assert np.allclose(svs[i][:k], f["singular_values"][i], rtol=1e-10) and np.all(svs[i][k:] == 0)
        t = A.contract(state[v1], state[v2])
        got = A.permute(t, sorted(t[1], key=repr)).ravel(order="F")
        assert rel(got, f["pair_products"][pair_off[i]:pair_off[i + 1]]) < 1e-9
        for ee in (int(e), ga.rev[e]):
            assert np.allclose(np.diag(new_msgs[ee])[:k], f["singular_values"][i], rtol=1e-10)


def test_host_compiled_device_code_reproduces_apply_fixture():
    from native_ctx import HostHarnessContext, build_hostlib

    f, ga, tensors, msgs, apply_state = _apply_fixture()
    ctx = HostHarnessContext(build_hostlib())
    ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
    ctx.set_dims(np.complex128, "norm", [2] * ga.nv, [3] * ga.ne)
    ctx.set_site_tensors(tensors)
    ctx.set_messages(msgs)
    svs = ctx.apply_two_site_gates([int(e) for e in f["edges"]], list(f["ops"]), max_rank=int(f["max_rank"]), normalize=True)
    shapes = [t.shape for t in tensors]
    new_tensors = [ctx.get_site_tensor(v).reshape(shapes[v], order="F") for v in range(ga.nv)]
    new_msgs = [m.reshape((3, 3), order="F") for m in ctx.msgs]
    _check_apply_result(f, ga, apply_state, new_tensors, new_msgs, svs)
    ctx.set_site_tensors(tensors)
    ctx.set_messages(msgs)
    v = int(f["one_site_vertex"])
    ctx.apply_one_site_gates([v], [f["one_site_op"]], normalize=True)
    assert rel(ctx.get_site_tensor(v), f["one_site_result"]) < 1e-12


def test_oracle_reproduces_apply_fixture(oracle):
    """The fixture's environment is what four oracle sweeps give, and the apply oracle regenerates its outputs."""
    from oracle import apply_oracle as A

    f, ga, tensors, msgs, apply_state = _apply_fixture()
    q = problems.synthetic_peps(graphs.named_grid((3, 3)), 3, 2, np.complex128, seed=int(f["seed"]))
    assert np.array_equal(np.concatenate([t.ravel(order="F") for t in q.tensors]), f["sites"])
    p = oracle.make_problem(ga, q.tensors, "norm")
    m = list(q.messages)
    for _ in range(4):
        m = oracle.sweep_jacobi(p, m)
    assert rel(np.concatenate([x.ravel(order="F") for x in m]), f["messages"]) < 1e-13
    state, env = apply_state(ga, tensors, msgs)
    for i, e in enumerate(f["edges"]):
        names = (("s", ga.src[e]), ("s", ga.dst[e]))
        _, new_env = A.apply_operator((f["ops"][i], names, names), state, env, trunc=int(f["max_rank"]), normalize=True)
        assert np.allclose(np.diag(new_env[(ga.src[e], ga.dst[e])]).real, f["singular_values"][i], rtol=1e-12)
