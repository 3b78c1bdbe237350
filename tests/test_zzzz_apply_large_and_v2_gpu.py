"""The largest gate shape (version 1) and the OPT-IN version 2 of the two-site gate kernel (csrc/bpx_apply2.cuh,
BPX_APPLY_V2=1) against version 1 on the GPU.  Neither has run on a GPU yet (host-verified: tests/test_apply_device_code.py);
this file sorts after every other test and bounds its own run time."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn
from itnn_b200 import graphs, problems
from test_zz_gpu_apply import bond_invariant, device_tensors, matching, oracle_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def test_gate_layer_at_the_true_cfg5_shape():
    """Degree-4 vertices with chi = 16, d = 2 (1 MiB tensors, 4096 x 32 matrix views): the four inner vertices of a 4 x 4
    lattice, two disjoint gates in one layer, against the apply oracle through random probes on the external legs (the
    full pair product would be 0.5 GB)."""
    from helpers import randn
    from itnn_b200 import graphs, problems
    from oracle import apply_oracle as A
    from test_zz_gpu_apply import device_tensors, oracle_state

    rng = np.random.default_rng(5)
    chi = 16
    p = problems.synthetic_peps(graphs.named_grid((4, 4)), chi, 2, np.float64, init="positive")
    ga = p.ga
    vid = {v: i for i, v in enumerate(ga.vertices)}
    edges = [ga.edge_index[(vid[(2, 2)], vid[(3, 2)])], ga.edge_index[(vid[(3, 3)], vid[(2, 3)])]]
    ops = [np.eye(4).reshape(2, 2, 2, 2) + 0.3 * randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(3, 0.0, True)
        msgs = ctx.get_messages()
        svs = ctx.apply_two_site_gates(edges, ops, max_rank=chi, normalize=True)
        new_tensors = device_tensors(ctx, p)
    state = oracle_state(p, p.tensors)
    env = {(ga.src[e], ga.dst[e]): msgs[e] for e in range(ga.ne)}

    def probe(t, slot, vecs):
        out = t
        for leg in reversed(range(t.ndim - 1)):
            if leg != slot:
                out = np.tensordot(out, vecs[leg], axes=([1 + leg], [0]))
        return out

    for e, op, sv in zip(edges, ops, svs):
        v1, v2, r = ga.src[e], ga.dst[e], ga.rev[e]
        assert new_tensors[v1].shape == (2, 16, 16, 16, 16) and new_tensors[v2].shape == (2, 16, 16, 16, 16)
        names = (("s", v1), ("s", v2))
        want_state, want_env = A.apply_operator((op, names, names), state, env, trunc=chi, normalize=True)
        assert np.allclose(sv, np.diag(want_env[(v1, v2)]).real, rtol=1e-8, atol=1e-12)
        w1 = A.permute(want_state[v1], state[v1][1])
        w2 = A.permute(want_state[v2], state[v2][1])
        p1 = [rng.standard_normal(chi) for _ in range(4)]
        p2 = [rng.standard_normal(chi) for _ in range(4)]
        got = probe(new_tensors[v1], ga.slot[e], p1) @ probe(new_tensors[v2], ga.slot[r], p2).T
        want = probe(w1, ga.slot[e], p1) @ probe(w2, ga.slot[r], p2).T
        assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("lattice,chi,max_rank,normalize", [((4, 4), 4, 0, False), ((3, 5), 3, 2, True), ((4, 4), 8, 8, True)])
def test_v2_matches_v1(monkeypatch, dtype, lattice, chi, max_rank, normalize):
    rng = np.random.default_rng(chi + max_rank)
    p = problems.synthetic_peps(graphs.named_grid(lattice), chi, 2, dtype, init="positive")
    edges = matching(p.ga, rng)
    ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
    results = []
    monkeypatch.setenv("BPX_APPLY_V3", "0")  # versions 1 / 2 are what version 3 hands its declined gates to
    for v2 in ("0", "1"):
        monkeypatch.setenv("BPX_APPLY_V2", v2)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(5, 0.0, True)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=normalize)
            assert ctx.apply_stats() == (0, 0)
            results.append((svs, device_tensors(ctx, p), ctx.get_messages()))
    (sv1, t1, m1), (sv2, t2, m2) = results
    s1, s2 = oracle_state(p, t1), oracle_state(p, t2)
    for e, a, b in zip(edges, sv1, sv2):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-14)
        v, w = p.ga.src[e], p.ga.dst[e]
        x, y = bond_invariant(s1, v, w), bond_invariant(s2, v, w)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(x).max()
    assert all(np.allclose(a, b, rtol=1e-10, atol=1e-14) for a, b in zip(m1, m2))


@pytest.mark.parametrize("v2", ["0", "1"])
def test_chunked_layers_equal_one_launch(monkeypatch, v2):
    """A work-space budget of 64 KiB splits the layer into several launches (the path cfg5-sized layers take when their
    work space exceeds half the free HBM): results must equal the single-launch run bit for bit."""
    rng = np.random.default_rng(2)
    p = problems.synthetic_peps(graphs.named_grid((4, 6)), 4, 2, np.float64, init="positive")
    edges = matching(p.ga, rng)
    ops = [randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    monkeypatch.setenv("BPX_APPLY_V2", v2)
    monkeypatch.setenv("BPX_APPLY_V3", "0")  # (version 3 has one work space per CTA and never chunks)
    out = []
    for budget in (None, str(64 * 1024)):
        if budget is None:
            monkeypatch.delenv("BPX_APPLY_WS_BYTES", raising=False)
        else:
            monkeypatch.setenv("BPX_APPLY_WS_BYTES", budget)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(4, 0.0, True)
            before = ctx.counters()["launches"]
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=3, normalize=True)
            out.append((svs, device_tensors(ctx, p), ctx.get_messages(), ctx.counters()["launches"] - before))
    (sv_a, t_a, m_a, n_a), (sv_b, t_b, m_b, n_b) = out
    assert n_a == 1 and n_b > 1
    assert all(np.array_equal(x, y) for x, y in zip(sv_a, sv_b))
    assert all(np.array_equal(x, y) for x, y in zip(t_a, t_b))
    assert all(np.array_equal(x, y) for x, y in zip(m_a, m_b))


# ---- version 3 (the Gram path, csrc/bpx_apply3.cuh): the default ------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("lattice,chi,max_rank,normalize", [((4, 4), 4, 0, False), ((3, 5), 3, 2, True), ((4, 4), 8, 8, True),
                                                           ((4, 4), 2, 0, True), ((2, 6), 5, 3, False)])
def test_v3_matches_v1(monkeypatch, dtype, lattice, chi, max_rank, normalize):
    """The default kernel against the step-by-step version 1 on a converged-ish BP environment: every gate of the layer is
    taken by the Gram path, singular values / new messages agree, tensors agree in the gauge-invariant pair product."""
    rng = np.random.default_rng(chi + max_rank)
    p = problems.synthetic_peps(graphs.named_grid(lattice), chi, 2, dtype, init="positive")
    edges = matching(p.ga, rng)
    ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
    results = []
    for v3 in ("0", "1"):
        monkeypatch.setenv("BPX_APPLY_V3", v3)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(5, 0.0, True)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=normalize)
            assert ctx.apply_stats() == ((len(edges), 0) if v3 == "1" else (0, 0))
            msgs = ctx.get_messages()
            # the update kernels see the new tensors (private images rebuilt); the bond gauge (signs / phases of the singular
            # vectors) differs between the kernels, so only gauge-invariant results of the next sweep are compared
            res, _ = ctx.sweep(1, 0.0, True)
            results.append((svs, device_tensors(ctx, p), msgs, (res, ctx.bethe_free_energy())))
    (sv1, t1, m1, r1), (sv3, t3, m3, r3) = results
    s1, s3 = oracle_state(p, t1), oracle_state(p, t3)
    for e, a, b in zip(edges, sv1, sv3):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-13)
        v, w = p.ga.src[e], p.ga.dst[e]
        x, y = bond_invariant(s1, v, w), bond_invariant(s3, v, w)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(x).max()
    assert all(np.allclose(a, b, rtol=1e-8, atol=1e-12) for a, b in zip(m1, m3))
    assert abs(r1[0] - r3[0]) <= 1e-9 and abs(r1[1] - r3[1]) <= 1e-8 * max(1.0, abs(r1[1]))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("lattice,chi,d,max_rank", [((3, 4), 3, 3, 0), ((4, 4), 5, 3, 4), ((3, 3), 2, 5, 0)])
def test_v3_matches_v1_with_an_odd_physical_dimension(monkeypatch, dtype, lattice, chi, d, max_rank):
    """d = 3 / 5: the matrix view has an odd number of columns (d chi), the absorb pass falls back to one column per batch and
    the last column group of the scratch copies is half empty -- the tensor-pipe Gram and final passes must mask it."""
    rng = np.random.default_rng(chi + d)
    p = problems.synthetic_peps(graphs.named_grid(lattice), chi, d, dtype, init="positive")
    edges = matching(p.ga, rng)
    ops = [np.eye(d * d).reshape(d, d, d, d) + 0.3 * randn(rng, dtype, (d, d, d, d)) for _ in edges]
    results = []
    for v3 in ("0", "1"):
        monkeypatch.setenv("BPX_APPLY_V3", v3)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(5, 0.0, True)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=True)
            taken, declined = ctx.apply_stats()
            assert (taken + declined == len(edges) and taken > 0) if v3 == "1" else (taken, declined) == (0, 0)
            res, _ = ctx.sweep(1, 0.0, True)
            results.append((svs, device_tensors(ctx, p), (res, ctx.bethe_free_energy())))
    (sv1, t1, r1), (sv3, t3, r3) = results
    s1, s3 = oracle_state(p, t1), oracle_state(p, t3)
    for e, a, b in zip(edges, sv1, sv3):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-13)
        v, w = p.ga.src[e], p.ga.dst[e]
        x, y = bond_invariant(s1, v, w), bond_invariant(s3, v, w)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(x).max()
    assert abs(r1[0] - r3[0]) <= 1e-9 and abs(r1[1] - r3[1]) <= 1e-8 * max(1.0, abs(r1[1]))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_v3_chunked_layer_is_bit_identical_to_one_chunk(monkeypatch, dtype):
    """The three gate kernels run once per CHUNK of the layer (per-gate work spaces, `status` indexed by the gate's position
    in the batch): a work-space budget that forces one, two or three gates per chunk -- with a last, smaller chunk -- must
    give exactly the tensors, messages and singular values of the single-chunk run."""
    rng = np.random.default_rng(21)
    p = problems.synthetic_peps(graphs.named_grid((5, 6)), 4, 2, dtype, init="positive")
    edges = matching(p.ga, rng)
    assert len(edges) >= 7
    ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
    out = []
    for budget in (None, "1", "300000", "450000"):   # bytes: default (one chunk), one gate per chunk, 2 - 6 gates per chunk
        if budget is None:
            monkeypatch.delenv("BPX_APPLY_WS_BYTES", raising=False)
        else:
            monkeypatch.setenv("BPX_APPLY_WS_BYTES", budget)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(4, 0.0, True)
            launches0 = ctx.counters()["launches"]
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=3, normalize=True)
            launches = ctx.counters()["launches"] - launches0
            assert ctx.apply_stats() == (len(edges), 0)
            out.append((svs, device_tensors(ctx, p), ctx.get_messages(), launches))
    sv0, t0, m0, l0 = out[0]
    assert l0 == 3                                     # sides, bond, final
    assert out[1][3] == 3 * len(edges)                 # one gate per chunk
    assert 3 < out[2][3] < 3 * len(edges) and 3 < out[3][3] < 3 * len(edges)
    for svs, ts, ms, _ in out[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(sv0, svs))
        assert all(np.array_equal(a, b) for a, b in zip(t0, ts))
        assert all(np.array_equal(a, b) for a, b in zip(m0, ms))


def test_v3_chunks_of_whole_waves_at_cfg5_shape_are_bit_identical(monkeypatch):
    """chi = 16, d = 2 on a 24 x 24 lattice (the bulk gates have cfg5's shape: 4096 x 32 matrix views, full TMA tiles): a
    work-space budget that splits the layer into chunks of 148 gates (whole waves of the side kernel) and a last smaller one
    gives exactly the singular values, messages and tensors of the single-chunk run."""
    rng = np.random.default_rng(4)
    p = problems.synthetic_peps(graphs.named_grid((24, 24)), 16, 2, np.float64, host_data=False)
    edges = matching(p.ga, rng)
    ops = [np.eye(4).reshape(2, 2, 2, 2) + 0.2 * randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    probe = sorted({int(p.ga.src[e]) for e in edges[::17]} | {int(p.ga.dst[e]) for e in edges[::17]})
    out = []
    for budget in (None, str(400 << 20)):
        if budget is None:
            monkeypatch.delenv("BPX_APPLY_WS_BYTES", raising=False)
        else:
            monkeypatch.setenv("BPX_APPLY_WS_BYTES", budget)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(3, 0.0, True)
            launches0 = ctx.counters()["launches"]
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=16, normalize=True)
            launches = ctx.counters()["launches"] - launches0
            assert ctx.apply_stats() == (len(edges), 0)
            res, _ = ctx.sweep(1, 0.0, True)
            out.append((np.stack(svs), [ctx.get_site_tensor(v) for v in probe], ctx.get_messages_flat(), res, launches))
    (sv0, t0, m0, r0, l0), (sv1, t1, m1, r1, l1) = out
    assert l0 == 3 and l1 == 3 * -(-len(edges) // 148), (l0, l1, len(edges))
    assert np.array_equal(sv0, sv1) and np.array_equal(m0, m1) and r0 == r1
    assert all(np.array_equal(a, b) for a, b in zip(t0, t1))
    assert np.all(sv0[:, 0] > 0) and np.allclose((sv0 ** 2).sum(axis=1), 1.0, rtol=1e-12)   # normalised singular values


def test_v3_declined_gates_fall_back_to_v1_bit_for_bit(monkeypatch):
    """An all-ones environment (test/test_apply_operator.jl:72) is rank one: the reference projects on the messages'
    support, the Gram path declines every gate, and the result is exactly what version 1 alone produces; mixed layers
    (some messages replaced by a full-rank one) split between the two kernels."""
    rng = np.random.default_rng(8)
    p = problems.synthetic_peps(graphs.named_grid((4, 4)), 3, 2, np.float64, init="ones")
    edges = matching(p.ga, rng)
    ops = [randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    out = []
    for v3 in ("0", "1"):
        monkeypatch.setenv("BPX_APPLY_V3", v3)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=3, normalize=True)
            if v3 == "1":
                assert ctx.apply_stats() == (0, len(edges))
            out.append((svs, device_tensors(ctx, p), ctx.get_messages()))
    (sv_a, t_a, m_a), (sv_b, t_b, m_b) = out
    assert all(np.array_equal(x, y) for x, y in zip(sv_a, sv_b))
    assert all(np.array_equal(x, y) for x, y in zip(t_a, t_b))
    assert all(np.array_equal(x, y) for x, y in zip(m_a, m_b))


# ---- two-site expectation values (csrc/bpx_expect2.cuh), never run on a GPU yet ------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_edge_expect_matches_oracle_gpu(oracle, dtype):
    """bpx_edge_expect on every directed edge of a 4 x 4 PEPS (chi = 4) after a few sweeps, against
    oracle.two_site_expect; the two orientations of an edge with the operator transposed accordingly agree."""
    rng = np.random.default_rng(4)
    p = problems.synthetic_peps(graphs.named_grid((4, 4)), 4, 2, dtype, init="positive")
    ga = p.ga
    edges = list(range(ga.ne))
    ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(5, 0.0, True)
        msgs = ctx.get_messages()
        num, den = ctx.edge_expect(edges, ops)
        swapped = [np.transpose(ops[e], (1, 0, 3, 2)) for e in edges]
        num_r, den_r = ctx.edge_expect([ga.rev[e] for e in edges], swapped)
        with pytest.raises(B.BPXError, match="out of range"):
            ctx.edge_expect([ga.ne], [ops[0]])
    op_ = oracle.make_problem(ga, p.tensors, "norm")
    for e in edges:
        want_num, want_den = oracle.two_site_expect(op_, msgs, e, ops[e])
        assert np.isclose(num[e], want_num, rtol=1e-10, atol=1e-13 * abs(want_den))
        assert np.isclose(den[e], want_den, rtol=1e-10)
    assert np.allclose(num_r, num, rtol=1e-10, atol=1e-13 * np.abs(den).max()) and np.allclose(den_r, den, rtol=1e-10)


@pytest.mark.parametrize("graph", ["chain", "comb"])
def test_energy_from_two_site_expectations_gpu(graph):
    from test_zzz_resident_state import check_energy_from_two_site_expectations

    check_energy_from_two_site_expectations(graphs.named_path_graph(5) if graph == "chain" else graphs.named_comb_tree((3, 2)))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_mirror_expect_two_site_gpu(oracle, dtype):
    from test_zz_apply_mirror import check_expect_two_site

    check_expect_two_site(oracle, dtype)
