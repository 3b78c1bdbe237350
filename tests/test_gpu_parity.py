"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): per-sweep messages within 1e-10 relative; converged local expectation
values within 1e-9.  Every test here needs a B200 and is marked `gpu`."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import peps_tensors, positive_messages, randn, rel_err, single_layer_tensors, spin_ice_tensors
from itnn_b200 import _lib, graphs, problems

pytestmark = pytest.mark.gpu

MSG_RTOL = 1e-10
EXPECT_ATOL = 1e-9
KERNELS = [_lib.BPX_KERNEL_AUTO, _lib.BPX_KERNEL_GENERIC]


def make_ctx(ga, dtype, mode, phys_dim, link_dim, tensors, msgs, kernel=_lib.BPX_KERNEL_AUTO):
    ctx = B.BPXContext(0)
    ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
    ctx.set_kernel_policy(kernel)
    ctx.set_dims(dtype, mode, phys_dim, link_dim)
    ctx.set_site_tensors(tensors)
    ctx.set_messages(msgs)
    return ctx


def check_sweeps(oracle, ga, dtype, mode, phys_dim, link_dim, tensors, msgs, nsweeps=3, kernel=_lib.BPX_KERNEL_AUTO,
                 normalize=True):
    p = oracle.make_problem(ga, tensors, mode)
    with make_ctx(ga, dtype, mode, phys_dim, link_dim, tensors, msgs, kernel) as ctx:
        want = list(msgs)
        for k in range(nsweeps):
            prev, want = want, oracle.sweep_jacobi(p, want, normalize)
            res, done = ctx.sweep(1, 0.0, normalize)
            got = ctx.get_messages()
            assert done == 1
            assert rel_err(got, want) < MSG_RTOL, f"sweep {k}"
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
        return ctx.buckets()


# ---- BASELINE configs ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("init", ["positive", "ones"])
def test_cfg1_4x4_chi2(oracle, kernel, init):
    p = problems.make_config("cfg1", init=init)
    check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 4, kernel)


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg2_32x32_chi8(oracle, kernel):
    p = problems.make_config("cfg2")
    check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 2, kernel)


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg3_heavy_hex_chi16_complex(oracle, kernel):
    p = problems.make_config("cfg3")
    buckets = check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 2, kernel)
    assert sorted(b["degree"] for b in buckets) == [1, 2, 3]


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg4_cubic_chi4_reduced(oracle, kernel):
    # same bucket as cfg4 (degree 6, chi 4, d 2) on a 4x4x4 periodic lattice so the oracle stays fast
    g = graphs.named_grid((4, 4, 4), periodic=True)
    p = problems.make_config("cfg4", graph=g)
    check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 2, kernel)


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg5_bucket_chi16_reduced(oracle, kernel):
    # same buckets as cfg5 (degrees 2/3/4, chi 16, d 2) on a 5x5 lattice
    g = graphs.named_grid((5, 5))
    p = problems.make_config("cfg5", graph=g)
    check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 2, kernel)


def test_complex_chi16_kernel_mixed_physical_dims(oracle):
    # the ComplexF64 chi = 16 on-chip kernel streams one physical slice at a time: d = 1, 2, 3 and degrees 1..3 together
    g = graphs.named_comb_tree((3, 2))
    g.add_edge((1, 2), (2, 2))  # a loop, and degree-3 vertices
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(11)
    phys = [1 + (v % 3) for v in range(ga.nv)]
    link_dim = [16] * ga.ne
    tensors = []
    for v in range(ga.nv):
        z = ga.row_ptr[v + 1] - ga.row_ptr[v]
        tensors.append(randn(rng, np.complex128, (phys[v], *([16] * z))) / np.sqrt(16.0**z))
    msgs = positive_messages(ga, link_dim, np.complex128, rng)
    buckets = check_sweeps(oracle, ga, np.complex128, "norm", phys, link_dim, tensors, msgs, 3)
    assert {b["degree"] for b in buckets} == {1, 2, 3}
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets)
    check_sweeps(oracle, ga, np.complex128, "norm", phys, link_dim, tensors, msgs, 2, normalize=False)


@pytest.mark.parametrize("dims,phys", [((5, 6), "uniform2"), ((4, 4), "mixed"), ((13, 13), "uniform2")])
def test_complex_chi8_kernel_square_lattice(oracle, dims, phys):
    # complex PEPS on a square lattice (the ComplexF64 twin of cfg2): degrees 2, 3, 4 in one launch of the complex
    # chi = 8 kernel; (13, 13) gives 121 degree-4 vertices = 242 half items > 148 CTAs (several rounds per CTA)
    ga = graphs.graph_arrays(graphs.named_grid(dims))
    rng = np.random.default_rng(5)
    physd = [2] * ga.nv if phys == "uniform2" else [1 + (v % 3) for v in range(ga.nv)]
    link_dim = [8] * ga.ne
    tensors = []
    for v in range(ga.nv):
        z = ga.row_ptr[v + 1] - ga.row_ptr[v]
        tensors.append(randn(rng, np.complex128, (physd[v], *([8] * z))) / np.sqrt(8.0**z))
    msgs = positive_messages(ga, link_dim, np.complex128, rng)
    buckets = check_sweeps(oracle, ga, np.complex128, "norm", physd, link_dim, tensors, msgs, 3 if dims != (13, 13) else 2)
    assert {b["degree"] for b in buckets} == {2, 3, 4}
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets)
    if dims == (5, 6):
        check_sweeps(oracle, ga, np.complex128, "norm", physd, link_dim, tensors, msgs, 2, normalize=False)
        # and the streamed host step (pinned buffers) on the same kernel
        import torch

        p = oracle.make_problem(ga, tensors, "norm")
        with make_ctx(ga, np.complex128, "norm", physd, link_dim, tensors, msgs) as ctx:
            flat = ctx.pack_messages(msgs)
            ta, tb = torch.empty(flat.size, dtype=torch.complex128).pin_memory(), torch.empty(flat.size, dtype=torch.complex128).pin_memory()
            a, b = ta.numpy(), tb.numpy()
            a[:] = flat
            want = list(msgs)
            for _ in range(3):
                prev, want = want, oracle.sweep_jacobi(p, want)
                res = ctx.sweep_host(a, b)
                assert rel_err(ctx.unpack_messages(b), want) < MSG_RTOL
                assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
                a, b = b, a


@pytest.mark.parametrize("chi,d,dims", [(2, 2, (4, 4)), (5, 3, (4, 5)), (7, 1, (3, 4)), (8, 3, (3, 3))])
def test_real_slice_kernel_small_and_odd_shapes(oracle, chi, d, dims):
    # Float64 buckets outside the hand-tuned (chi = 8, d = 2) shape -- cfg1's chi = 2, odd d, chi = 5 / 7 -- run on the
    # templated slice kernel (zero-padded image, physical pairs): degrees 2, 3, 4 in one launch
    ga = graphs.graph_arrays(graphs.named_grid(dims))
    rng = np.random.default_rng(chi * 10 + d)
    link_dim = [chi] * ga.ne
    tensors = peps_tensors(ga, chi, d, np.float64, rng)
    msgs = positive_messages(ga, link_dim, np.float64, rng)
    buckets = check_sweeps(oracle, ga, np.float64, "norm", [d] * ga.nv, link_dim, tensors, msgs, 3)
    assert {b["degree"] for b in buckets} == {2, 3, 4}
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets), buckets
    check_sweeps(oracle, ga, np.float64, "norm", [d] * ga.nv, link_dim, tensors, msgs, 2, normalize=False)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("chi,d", [(12, 3), (16, 1), (9, 2)])
def test_wide_slice_kernel_dims_up_to_16(oracle, dtype, chi, d):
    # degree 1..3 with link dims 9..16 (and chi = 16 outside the tuned d = 2 shape): the 16-wide slice kernel, both dtypes
    g = graphs.named_comb_tree((3, 2))
    g.add_edge((1, 2), (2, 2))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(chi + d)
    link_dim = [chi] * ga.ne
    tensors = peps_tensors(ga, chi, d, dtype, rng)
    msgs = positive_messages(ga, link_dim, dtype, rng)
    buckets = check_sweeps(oracle, ga, dtype, "norm", [d] * ga.nv, link_dim, tensors, msgs, 3)
    assert {b["degree"] for b in buckets} == {1, 2, 3}
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets), buckets


# ---- edge cases ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_ragged_link_dims_and_degree_one(oracle, dtype):
    g = graphs.named_comb_tree((3, 3))
    g.add_edge((1, 3), (2, 3))  # one loop
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(3)
    link_dim = [0] * ga.ne
    for e in range(ga.ne):
        link_dim[e] = link_dim[ga.rev[e]] = 1 + (min(e, ga.rev[e]) % 4)  # dims 1..4, including 1
    phys = [1 + (v % 3) for v in range(ga.nv)]
    tensors = []
    for v in range(ga.nv):
        dims = [link_dim[f] for f in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
        tensors.append(randn(rng, dtype, (phys[v], *dims)))
    msgs = positive_messages(ga, link_dim, dtype, rng)
    buckets = check_sweeps(oracle, ga, dtype, "norm", phys, link_dim, tensors, msgs, 3)
    # link dims 1..4, different per leg, d = 1..3, degrees 1..3: all on the slice kernels (zero-padded private images)
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets), buckets
    check_sweeps(oracle, ga, dtype, "norm", phys, link_dim, tensors, msgs, 2, kernel=_lib.BPX_KERNEL_GENERIC)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_ragged_link_dims_9_to_16_on_degree_3_graphs(oracle, dtype):
    """Per-leg different link dims in 9..16 on vertices of degree 1..3 (no degree-4 vertex, so nothing is padded by
    bpx_pad.cuh): the 16-wide slice kernels mask message fragments and stores to the true dims."""
    g = graphs.named_comb_tree((3, 3))
    g.add_edge((1, 3), (2, 3))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(4)
    link_dim = [0] * ga.ne
    for e in range(ga.ne):
        link_dim[e] = link_dim[ga.rev[e]] = 9 + (min(e, ga.rev[e]) % 8)
    phys = [2] * ga.nv
    tensors = []
    for v in range(ga.nv):
        dims = [link_dim[f] for f in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
        tensors.append(randn(rng, dtype, (2, *dims)) / np.sqrt(np.prod(dims)))
    msgs = positive_messages(ga, link_dim, dtype, rng)
    buckets = check_sweeps(oracle, ga, dtype, "norm", phys, link_dim, tensors, msgs, 3)
    assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in buckets), buckets


def test_isolated_vertex_and_empty_graph(oracle):
    with B.BPXContext(0) as ctx:
        ctx.set_graph([], [], [], 0)
        ctx.set_dims(np.float64, "norm", [], [])
        res, done = ctx.sweep(2, 0.0)
        assert done == 2
    g = graphs.NamedGraph([0, 1, 2])
    g.add_edge(0, 1)  # vertex 2 isolated
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(0)
    tensors = [randn(rng, np.float64, (2, 3)), randn(rng, np.float64, (2, 3)), randn(rng, np.float64, (2,))]
    msgs = positive_messages(ga, [3] * ga.ne, np.float64, rng)
    check_sweeps(oracle, ga, np.float64, "norm", [2, 2, 2], [3] * ga.ne, tensors, msgs, 1)
    p = oracle.make_problem(ga, tensors, "norm")
    with make_ctx(ga, np.float64, "norm", [2, 2, 2], [3, 3], tensors, msgs) as ctx:
        assert np.allclose(ctx.vertex_scalars(), oracle.vertex_scalars(p, msgs), rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_zero_sum_guard_and_no_normalize(oracle, dtype):
    # antisymmetric site tensors make some raw messages sum to exactly zero -> left unnormalised
    g = graphs.named_path_graph(3)
    ga = graphs.graph_arrays(g)
    A0 = np.zeros((2, 2), dtype=dtype)
    A0[0, 0], A0[1, 1] = 1.0, 1.0
    A1 = np.zeros((2, 2, 2), dtype=dtype)
    A1[0, 0, 1], A1[0, 1, 0] = 1.0, -1.0
    A1[1, 0, 1], A1[1, 1, 0] = 1.0, 1.0
    tensors = [A0, A1, A0.copy()]
    msgs = [np.array([[1.0, -1.0], [1.0, -1.0]], dtype=dtype) for _ in range(ga.ne)]  # sum == 0 exactly
    p = oracle.make_problem(ga, tensors, "norm")
    want = oracle.sweep_jacobi(p, msgs)
    with make_ctx(ga, dtype, "norm", [2] * 3, [2] * ga.ne, tensors, msgs) as ctx:
        ctx.sweep(1)
        got = ctx.get_messages()
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-15)
    assert any(abs(w.sum()) < 1e-300 for w in want), "the case must exercise the iszero branch"
    rng = np.random.default_rng(1)
    tensors = peps_tensors(ga, 2, 2, dtype, rng)
    msgs = positive_messages(ga, [2] * ga.ne, dtype, rng)
    check_sweeps(oracle, ga, dtype, "norm", [2] * 3, [2] * ga.ne, tensors, msgs, 2, normalize=False)


# ---- reference known answers end to end on the GPU (sequential schedule + single-layer mode) ----------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_tree_exact_one_sequential_sweep_gpu(oracle, dtype):
    # test/test_beliefpropagation.jl:157-202 through the reference-shaped API
    for g, chi in ((graphs.named_grid((2, 1)), 2), (graphs.named_comb_tree((4, 3)), 3)):
        rng = np.random.default_rng(123)
        links = {frozenset((e.src, e.dst)): B.Index(chi) for e in g.edges()}
        tn = B.tensornetwork(lambda v: B.randn_itensor(rng, dtype, [links[frozenset((e.src, e.dst))] for e in g.incident_edges(v)]),
                             g.vertices())
        messages = {e: B.ITensor(np.ones(chi, dtype=dtype), (tn.linkind(e),)) for e in g.all_edges()}
        cache = B.beliefpropagation(tn, messages, stopping_criterion=dict(maxiter=1))
        z_bp = np.exp(B.bethe_free_energy(tn, cache))
        cp = B.canonical_arrays(tn)
        z_exact = oracle.contract_all(oracle.make_problem(cp.ga, cp.tensors, "single"))
        assert np.isclose(z_bp, z_exact, rtol=np.finfo(np.float64).eps ** (1 / 3))


@pytest.mark.parametrize("n", [3, 4, 5])
def test_spin_ice_gpu(oracle, n):
    # test/test_beliefpropagation.jl:204-225
    g = graphs.named_grid((n, n), periodic=True)
    rng = np.random.default_rng(123)
    links = {frozenset((e.src, e.dst)): B.Index(2) for e in g.edges()}
    ga0 = graphs.graph_arrays(g)
    t = spin_ice_tensors(ga0)[0]
    tn = B.tensornetwork(lambda v: B.ITensor(t, [links[frozenset((e.src, e.dst))] for e in g.incident_edges(v)]), g.vertices())
    messages = {e: B.ITensor(rng.random(2), (tn.linkind(e),)) for e in g.all_edges()}
    info = B.BeliefPropagationResult()
    cache = B.beliefpropagation(tn, messages, stopping_criterion=dict(maxiter=10, tol=1e-10), info=info)
    assert np.isclose(np.exp(B.bethe_free_energy(tn, cache)), 1.5 ** (n * n))
    assert 1 <= info.iterations <= 10


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_sequential_schedule_matches_oracle(oracle, dtype):
    g = graphs.named_grid((3, 4))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(9)
    tensors = peps_tensors(ga, 3, 2, dtype, rng)
    msgs = positive_messages(ga, [3] * ga.ne, dtype, rng)
    seq = [ga.edge_id(e) for e in graphs.forest_cover_edge_sequence(g)]
    p = oracle.make_problem(ga, tensors, "norm")
    with make_ctx(ga, dtype, "norm", [2] * ga.nv, [3] * ga.ne, tensors, msgs) as ctx:
        want = list(msgs)
        for _ in range(3):
            prev, want = want, oracle.sweep_sequential(p, want, seq)
            res, done = ctx.sweep_sequence(seq, 1)
            assert rel_err(ctx.get_messages(), want) < MSG_RTOL
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
        # a sequence that repeats an edge and skips others is still honoured literally
        odd = [seq[0], seq[5], seq[0], seq[7]]
        want = oracle.sweep_sequential(p, want, odd)
        ctx.sweep_sequence(odd, 1)
        assert rel_err(ctx.get_messages(), want) < MSG_RTOL


# ---- NormNetwork BP through the reference-shaped API: test/test_apply_operator.jl:62-74 ---------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("gname", ["cycle4", "path4"])
def test_normnetwork_bp_api_and_expectation_values(oracle, dtype, gname):
    g = graphs.named_cycle_graph(4) if gname == "cycle4" else graphs.named_path_graph(4)
    tn, _, _ = B.random_state(dtype, g, d=3, chi=3, rng=np.random.default_rng(123))
    nn = B.normnetwork(tn)
    env0 = B.message_environment(B.ones_message, nn)
    info = B.BeliefPropagationResult()
    env = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=100, tol=1e-13), info=info)
    cp = B.canonical_arrays(nn)
    p = oracle.make_problem(cp.ga, cp.tensors, "norm")
    seq = [cp.ga.edge_id(e) for e in B.default_beliefpropagation_edges(nn)]
    msgs0 = [np.ones((3, 3), dtype=dtype) for _ in range(cp.ga.ne)]
    want, it, delta = oracle.beliefpropagation(p, msgs0, maxiter=100, tol=1e-13, schedule="sequential", edge_seq=seq)
    assert info.iterations == it
    got = [env[cp.ga.named_edge(e)].array(cp.bra_names[e], cp.ket_names[e]) for e in range(cp.ga.ne)]
    assert rel_err(got, want) < 1e-9
    # synchronous strategy: same fixed point -> same local expectation values (1e-9)
    env_sync = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=500, tol=1e-15),
                                   message_update_algorithm=B.B200MessageUpdate())
    sz = np.diag([1.0, 0.0, -1.0])
    e_gpu = np.array(B.expect(nn, env_sync, sz))
    e_ref = np.array([oracle.local_expect(p, want, v, sz.astype(dtype)) for v in range(cp.ga.nv)])
    assert np.allclose(e_gpu, e_ref, atol=EXPECT_ATOL)
    assert np.allclose(B.vertex_scalars(nn, env), oracle.vertex_scalars(p, want), rtol=1e-9)
    assert np.allclose(B.edge_scalars(env), oracle.edge_scalars(p, want), rtol=1e-9)
    if gname == "path4":  # tree: BP is exact
        z = np.exp(B.bethe_free_energy(nn, env))
        assert np.isclose(z, oracle.contract_all(p).real, rtol=1e-9)


def test_converged_expectation_values_cfg1(oracle):
    p = problems.make_config("cfg1")
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    want, it, delta = oracle.beliefpropagation(op, p.messages, maxiter=300, tol=1e-14)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        res, done = ctx.sweep(300, 1e-14)
        assert done == it and res < 1e-14
        hist = ctx.residual_history()
        assert len(hist) == done and hist[-1] == res
        sz = np.diag([1.0, -1.0])
        num = ctx.vertex_expect_numerators([sz] * p.ga.nv)
        den = ctx.vertex_scalars()
        e_ref = [oracle.local_expect(op, want, v, sz) for v in range(p.ga.nv)]
        assert np.allclose(num / den, e_ref, atol=EXPECT_ATOL)
        assert abs(ctx.iterate_diff(want)) < 1e-12
        assert abs(ctx.iterate_diff(p.messages) - oracle.iterate_diff(want, p.messages)) < 1e-11


def test_message_update_single_edge_and_iterate_diff(oracle):
    g = graphs.named_grid((3, 3))
    tn, _, _ = B.random_state(np.float64, g, d=2, chi=2, rng=np.random.default_rng(4))
    nn = B.normnetwork(tn)
    cache = B.message_environment(B.identity_message, nn)
    before = cache.copy()
    e = graphs.NamedEdge((2, 2), (2, 3))
    B.message_update(cache, nn, e)
    cp = B.canonical_arrays(nn)
    p = oracle.make_problem(cp.ga, cp.tensors, "norm")
    want = oracle.message_update(p, [np.eye(2) for _ in range(cp.ga.ne)], cp.ga.edge_id(e))
    assert np.allclose(cache[e].array(cp.bra_names[cp.ga.edge_id(e)], cp.ket_names[cp.ga.edge_id(e)]), want, rtol=1e-12)
    d = B.iterate_diff(cache, before)
    msgs_after = [np.eye(2) for _ in range(cp.ga.ne)]
    msgs_after[cp.ga.edge_id(e)] = want
    assert abs(d - oracle.iterate_diff(msgs_after, [np.eye(2)] * cp.ga.ne)) < 1e-12


def test_device_side_synthetic_inputs_match_host_recipe(oracle):
    # bpx_fill_synthetic must reproduce problems.synthetic_peps (up to the last ulp of device log/cos)
    g = graphs.named_grid((3, 3))
    p = problems.make_config("cfg2", graph=g)
    q = problems.make_config("cfg2", graph=g, host_data=False)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, q)
        msgs = ctx.get_messages()
        for e in range(p.ga.ne):
            assert np.allclose(msgs[e], p.messages[e], rtol=1e-12, atol=1e-15)
        for v in (0, 4, 8):
            assert np.allclose(ctx.get_site_tensor(v), p.tensors[v].ravel(order="F"), rtol=1e-11, atol=1e-14)
        res, done = ctx.sweep(2)
        got = ctx.get_messages()
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    want = oracle.sweep_jacobi(op, oracle.sweep_jacobi(op, p.messages))
    assert rel_err(got, want) < 1e-9


def test_sweep_host_single_call_step(oracle):
    # the e2e entry point: host iterate in, one sweep, host iterate + residual out
    p = problems.make_config("cfg2", graph=graphs.named_grid((6, 6)))
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        a = ctx.pack_messages(p.messages)
        b = np.empty_like(a)
        want = list(p.messages)
        for _ in range(3):
            prev, want = want, oracle.sweep_jacobi(op, want)
            res = ctx.sweep_host(a, b)
            assert rel_err(ctx.unpack_messages(b), want) < MSG_RTOL
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
            a, b = b, a


@pytest.mark.parametrize("name,dims,periodic", [("cfg2", (6, 6), False), ("cfg2", (32, 32), False), ("cfg4", (4, 4, 4), True),
                                                ("cfg3", None, False), ("cfg5", (5, 5), False), ("cfg1", None, False)])
def test_sweep_host_streamed_io_with_pinned_buffers(oracle, name, dims, periodic):
    # pinned host buffers: the sweep kernel runs while the upload arrives in chunks (items gated on a progress word)
    # and stores the new messages straight into the caller's buffer; results must equal the staged path and the oracle
    # (cfg5 / cfg1: sweeps of several launches or generic buckets take the staged path even with pinned buffers)
    import torch

    p = problems.make_config(name, graph=graphs.named_grid(dims, periodic=periodic) if dims else None)
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        flat = ctx.pack_messages(p.messages)
        tdt = torch.complex128 if flat.dtype.kind == "c" else torch.float64
        pa, pb = torch.empty(flat.size, dtype=tdt).pin_memory(), torch.empty(flat.size, dtype=tdt).pin_memory()
        a, b = pa.numpy(), pb.numpy()
        a[:] = flat
        b[:] = np.nan
        staged_in, staged_out = flat.copy(), np.empty_like(flat)
        want = list(p.messages)
        for _ in range(3):
            prev, want = want, oracle.sweep_jacobi(op, want)
            res = ctx.sweep_host(a, b)
            assert np.array_equal(ctx.get_messages_flat(), b)  # the device copy of the iterate is complete too
            res_staged = ctx.sweep_host(staged_in, staged_out)  # pageable memory: upload, sweep, download
            assert np.array_equal(b, staged_out) and res == res_staged
            assert rel_err(ctx.unpack_messages(b), want) < MSG_RTOL
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
            a, b = b, a
            staged_in, staged_out = staged_out, staged_in


def test_sweep_host_streams_through_registered_numpy_buffers(oracle):
    # what the Julia glue does (no CUDA binding of its own): bpx_host_register on ordinary arrays
    p = problems.make_config("cfg2", graph=graphs.named_grid((7, 5)))
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        a = ctx.pack_messages(p.messages)
        b = np.full_like(a, np.nan)
        ctx.host_register(a)
        ctx.host_register(b)
        try:
            want = list(p.messages)
            for _ in range(4):  # the second use of a buffer pair replays the cached graph
                prev, want = want, oracle.sweep_jacobi(op, want)
                res = ctx.sweep_host(a, b)
                assert rel_err(ctx.unpack_messages(b), want) < MSG_RTOL
                assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
                assert abs(ctx.last_residual() - res) < 1e-15
                a, b = b, a
        finally:
            ctx.host_unregister(a)
            ctx.host_unregister(b)
        res, done = ctx.sweep(1, 0.0)  # ordinary sweeps after streamed steps (residual ring re-used slots)
        prev, want = want, oracle.sweep_jacobi(op, want)
        assert done == 1 and abs(res - oracle.iterate_diff(want, prev)) < 1e-11
        assert rel_err(ctx.get_messages(), want) < MSG_RTOL


def test_streamed_step_falls_back_to_the_staged_path_when_the_upload_stalls(oracle, monkeypatch):
    # something serialises kernel and copies (a profiler replaying kernels): the kernel gives up waiting after ~4 s, the
    # step is repeated staged -- same results -- and the context stops streaming
    import torch

    p = problems.make_config("cfg2", graph=graphs.named_grid((5, 5)))
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        flat = ctx.pack_messages(p.messages)
        ta, tb = torch.empty(flat.size, dtype=torch.float64).pin_memory(), torch.empty(flat.size, dtype=torch.float64).pin_memory()
        a, b = ta.numpy(), tb.numpy()
        a[:] = flat
        want = oracle.sweep_jacobi(op, p.messages)
        monkeypatch.setenv("BPX_IO_TEST_STALL", "1")  # the upload never reports progress
        res = ctx.sweep_host(a, b)
        monkeypatch.delenv("BPX_IO_TEST_STALL")
        assert rel_err(ctx.unpack_messages(b), want) < MSG_RTOL
        assert abs(res - oracle.iterate_diff(want, p.messages)) < 1e-11
        prev, want = want, oracle.sweep_jacobi(op, want)
        res = ctx.sweep_host(b, a)  # staged from now on
        assert rel_err(ctx.unpack_messages(a), want) < MSG_RTOL
        assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11


@pytest.mark.parametrize("name,dims", [("cfg5", (10, 10)), ("cfg2", (32, 32)), ("cfg4", (16, 16, 16)), ("cfg5", (64, 64))])
def test_full_path_sampled_edges_against_oracle(oracle, name, dims):
    """Size-independent check of the production path (device-generated inputs, specialised kernels): recompute a
    random sample of directed edges with the oracle from the GPU's own inputs; plus the sum-normalisation property."""
    g = graphs.named_grid(dims, periodic=(name == "cfg4"))  # cfg4 at its full BASELINE size (16^3 periodic); cfg5: 64x64 of 256x256
    q = problems.make_config(name, graph=g, host_data=False)
    rng = np.random.default_rng(1)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, q)
        before = ctx.get_messages()
        res, done = ctx.sweep(1)
        after = ctx.get_messages()
        kernels = {b["degree"]: b["kernel"] for b in ctx.buckets()}
        assert kernels[max(kernels)] in (_lib.BPX_KERNEL_ONCHIP, _lib.BPX_KERNEL_SLICED)
        ga = q.ga
        sample = rng.choice(ga.ne, size=48, replace=False)
        worst = 0.0
        for e in sample:
            u = ga.src[e]
            z = ga.row_ptr[u + 1] - ga.row_ptr[u]
            A = ctx.get_site_tensor(u).reshape((q.d,) + (q.chi,) * z, order="F")
            ins = [None if f == e else before[ga.rev[f]] for f in range(ga.row_ptr[u], ga.row_ptr[u + 1])]
            want = oracle.normalize_message(oracle.contract_norm(A, ga.slot[e], ins))
            worst = max(worst, np.abs(after[e] - want).max() / np.abs(want).max())
        assert worst < MSG_RTOL, worst
        sums = np.array([m.sum() for m in after])
        assert np.allclose(sums, 1.0, rtol=0, atol=1e-12)
        assert abs(res - oracle.iterate_diff(after, before)) < 1e-11


# ---- hardening of the specialised kernels -------------------------------------------------------------
@pytest.mark.parametrize("name,dims", [("cfg2", (5, 6)), ("cfg5", (4, 4)), ("cfg4", (4, 4, 4))])
def test_fast_kernels_without_normalisation_and_multi_sweep(oracle, name, dims):
    g = graphs.named_grid(dims, periodic=(name == "cfg4"))
    p = problems.make_config(name, graph=g)
    # unnormalised messages grow quickly: two sweeps are enough to exercise the branch
    check_sweeps(oracle, p.ga, p.dtype, "norm", p.phys_dim, p.link_dim, p.tensors, p.messages, 2, normalize=False)


def test_fast_kernel_converges_like_the_oracle(oracle):
    g = graphs.named_grid((6, 6))
    p = problems.make_config("cfg2", graph=g)
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    hist_want = []
    want, it, delta = oracle.beliefpropagation(op, p.messages, maxiter=200, tol=1e-12, history=hist_want)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        assert all(b["kernel"] == _lib.BPX_KERNEL_ONCHIP for b in ctx.buckets())
        res, done = ctx.sweep(200, 1e-12)
        assert done == it
        hist = ctx.residual_history()
        assert len(hist) == done
        assert np.allclose(hist, hist_want, rtol=1e-6, atol=1e-13)
        assert rel_err(ctx.get_messages(), want) < 1e-9
        # async sweeps append to the same history and keep the iterate consistent
        ctx.sweep_async(3)
        assert len(ctx.residual_history()) == done + 3
        assert ctx.counters()["sweeps"] == done + 3


def test_nan_in_a_site_tensor_propagates_to_the_residual(oracle):
    # Julia's `maximum` propagates NaN (beliefpropagation.jl:262); so must the fused residual key
    g = graphs.named_grid((4, 4))
    p = problems.make_config("cfg2", graph=g)
    t = [x.copy() for x in p.tensors]
    t[5][0, 0, 0, 0, 0] = np.nan
    with B.BPXContext(0) as ctx:
        ctx.set_graph(p.ga.src, p.ga.dst, p.ga.slot, p.ga.nv)
        ctx.set_dims(p.dtype, "norm", p.phys_dim, p.link_dim)
        ctx.set_site_tensors(t)
        ctx.set_messages(p.messages)
        res, done = ctx.sweep(1)
        assert np.isnan(res)


def test_site_tensor_update_refreshes_private_images(oracle):
    # bpx_set_site_tensor after a sweep must invalidate the pre-swizzled image
    g = graphs.named_grid((4, 4))
    p = problems.make_config("cfg2", graph=g)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(1)
        t = [x.copy() for x in p.tensors]
        t[5] = t[5] * 2.0 + 0.01
        flat = np.ascontiguousarray(t[5].ravel(order="F"))
        import ctypes as C
        ctx._check(ctx.lib.bpx_set_site_tensor(ctx.h, 5, flat.ctypes.data_as(C.c_void_p)))
        ctx.set_messages(p.messages)
        ctx.sweep(1)
        got = ctx.get_messages()
    want = oracle.sweep_jacobi(oracle.make_problem(p.ga, t, "norm"), p.messages)
    assert rel_err(got, want) < MSG_RTOL


@pytest.mark.parametrize("variant", [{}, {"BPX_SLICED_CW": "16"}, {"BPX_SLICED_G": "4"}], ids=["g8-cw8", "g8-cw16", "g4-cw8"])
def test_group_cooperative_sliced_kernel_equals_the_independent_kernel_at_scale(monkeypatch, variant):
    """The group-cooperative chi = 16 kernel (bpx_sliced2.cuh) hands partial tiles between warps, CTAs and vertices through
    counters; a protocol hole shows up as a handful of wrong messages only when every group processes MANY vertices (a
    first version passed every small-lattice test and failed here).  48 x 48: 111 vertices per group; six sweeps must equal
    version 1 (one CTA per half vertex, no cross-CTA protocol) to rounding, and the run must converge like it."""
    g = graphs.named_grid((48, 48))
    q = problems.make_config("cfg5", graph=g, host_data=False)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = []
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, q)
            assert {b["degree"]: b["kernel"] for b in ctx.buckets()}[4] == _lib.BPX_KERNEL_SLICED
            for _ in range(6):
                res, _ = ctx.sweep(1)
                out.append((res, ctx.get_messages_flat()))
            res, done = ctx.sweep(200, 1e-10)
        for k in env:
            monkeypatch.delenv(k)
        return out, (res, done)

    want, conv_want = run({"BPX_SLICED_V1": "1"})
    got, conv_got = run(variant)
    for (ra, ma), (rb, mb) in zip(want, got):
        assert np.abs(ma - mb).max() < 1e-13
        assert abs(ra - rb) < 1e-12
    assert conv_got[1] == conv_want[1] and conv_got[0] < 1e-10
