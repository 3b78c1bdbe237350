"""Host-side mirror of the reference interface: containers, criteria, algorithm selection (CPU only)."""
import numpy as np
import pytest

import itnn_b200 as B
from itnn_b200 import graphs


def test_named_grid_matches_baseline_counts():
    g = graphs.named_grid((4, 4))
    assert (g.nv(), 2 * g.ne()) == (16, 48)
    degs = sorted(g.degree(v) for v in g.vertices())
    assert (degs.count(2), degs.count(3), degs.count(4)) == (4, 8, 4)
    g = graphs.named_grid((32, 32))
    assert (g.nv(), 2 * g.ne()) == (1024, 3968)
    g = graphs.named_grid((16, 16, 16), periodic=True)
    assert (g.nv(), 2 * g.ne()) == (4096, 24576) and all(g.degree(v) == 6 for v in g.vertices())
    g = graphs.heavy_hex_127()
    assert (g.nv(), 2 * g.ne()) == (127, 288)
    assert max(g.degree(v) for v in g.vertices()) == 3 and min(g.degree(v) for v in g.vertices()) == 1
    assert len(graphs.connected_components(g)) == 1


def test_graph_arrays_are_consistent():
    g = graphs.named_grid((3, 4), periodic=True)
    ga = graphs.graph_arrays(g)
    for e in range(ga.ne):
        r = ga.rev[e]
        assert (ga.src[r], ga.dst[r]) == (ga.dst[e], ga.src[e])
        assert ga.row_ptr[ga.src[e]] + ga.slot[e] == e


# -- MessageCache container semantics: test/test_beliefpropagation.jl:39-82 ---------------------------
def test_messagecache_basics():
    g = graphs.named_grid((3, 3))
    bpc = B.messagecache(lambda e: f"{e.src} => {e.dst}", g.all_edges())
    assert len(bpc) == 2 * g.ne()
    assert bpc[((1, 1), (1, 2))] == "(1, 1) => (1, 2)"
    bpc[((1, 1), (1, 2))] = "new message"
    assert bpc[((1, 1), (1, 2))] == "new message"
    pairs = {((1, 2), (2, 2)): "m1", ((2, 2), (2, 3)): "m2"}
    new = bpc.copy().copyto(pairs)
    assert new[((1, 1), (1, 2))] == "new message" and new[((1, 2), (2, 2))] == "m1" and new[((2, 2), (2, 3))] == "m2"
    dst = B.messagecache(lambda e: "", g.all_edges())
    dst.copyto(bpc, [((1, 2), (2, 2)), ((2, 2), (2, 3))])
    assert dst[((1, 1), (1, 2))] == "" and dst[((1, 2), (2, 2))] == "(1, 2) => (2, 2)"


def test_incoming_messages_exclusion_rule():  # test/test_beliefpropagation.jl:104-114
    g = graphs.named_path_graph(3)
    bpc = B.messagecache(lambda e: (e.src, e.dst), g.all_edges())
    assert B.incoming_messages(bpc, (2, 3)) == [(1, 2)]
    assert B.incoming_messages(bpc, (1, 2)) == []
    assert B.incoming_messages(bpc, (2, 1)) == [(3, 2)]


def test_subgraph():  # test/test_beliefpropagation.jl:117-133
    g = graphs.named_grid((3,))
    bpc = B.messagecache(lambda e: 1, g.all_edges())
    sub = bpc.subgraph([(1,), (2,)])
    assert set(sub.vertices()) == {(1,), (2,)} and sub.has_edge(((1,), (2,)))


# -- stopping criterion shorthand: beliefpropagation.jl:16-55 -----------------------------------------
def test_stopping_criterion_selection():
    sel = B.select_beliefpropagation_stopping_criterion
    assert sel(dict(maxiter=3)) == B.StopAfterIteration(3)
    assert sel(dict(tol=1e-8)) == B.StopWhenConverged(1e-8)
    c = sel(dict(maxiter=3, tol=1e-8))
    assert c.criteria == [B.StopAfterIteration(3), B.StopWhenConverged(1e-8)]
    explicit = B.StopAfterIteration(10) | B.StopWhenConverged(1e-10)
    assert sel(explicit) is explicit
    with pytest.raises(B.ArgumentError, match="must be specified"):
        sel(None)
    with pytest.raises(B.ArgumentError, match="Unrecognized"):
        sel(dict(maxiter=1, foo=2))
    with pytest.raises(B.ArgumentError, match="At least one"):
        sel(dict())


# -- select_algorithm: src/select_algorithm.jl:16-51 --------------------------------------------------
def test_select_algorithm():
    f = B.message_update
    assert isinstance(B.select_algorithm(f, None), B.SimpleMessageUpdate)
    assert B.select_algorithm(f, dict(normalize=False)).normalize is False
    inst = B.B200MessageUpdate(normalize=False)
    assert B.select_algorithm(f, inst) is inst
    with pytest.raises(B.ArgumentError):
        B.select_algorithm(f, inst, normalize=True)
    with pytest.raises(B.ArgumentError):
        B.select_algorithm(f, dict(normalize=True), normalize=True)
    with pytest.raises(TypeError):
        B.default_algorithm(print)


# -- NormNetwork name map and views: test/test_normnetwork.jl:60-135 ----------------------------------
def test_normnetwork_names_and_canonical_layout():
    g = graphs.named_path_graph(3)
    tn, l, s = B.random_state(np.float64, g, d=2, chi=3)
    nn = B.normnetwork(tn)
    assert set(nn.vertices()) == set(tn.vertices())
    e = graphs.NamedEdge(1, 2)
    kn = tn.linkname(e)
    assert nn.braname(kn) != kn and nn.braname(s[1].name) == s[1].name
    with pytest.raises(KeyError):
        nn.braname("nope")
    assert B.KetView(nn)[2] is tn[2]
    assert np.array_equal(B.BraView(nn)[2].data, np.conj(tn[2].data))
    assert B.BraView(nn).linkname(e) == nn.braname(kn)
    custom = B.normnetwork(tn, {n: ("bra", n) for n in tn.dimname_vertices})
    assert custom.braname(kn) == ("bra", kn)
    cp = B.canonical_arrays(nn)
    assert cp.mode == "norm" and cp.phys_dim == [2, 2, 2]
    assert [t.shape for t in cp.tensors] == [(2, 3), (2, 3, 3), (2, 3)]
    env = B.message_environment(B.ones_message, nn)
    m = env[e]
    assert m.dimnames() == (nn.braname(kn), kn) and np.all(m.data == 1)


def test_canonical_arrays_permutes_by_name():
    # a tensor stored as (link, site, link) must come out as [site, links in neighbour order]
    a, b, s1, s2, s3 = (B.Index(2, "a"), B.Index(3, "b"), B.Index(2, "s1"), B.Index(2, "s2"), B.Index(2, "s3"))
    rng = np.random.default_rng(0)
    t1 = B.randn_itensor(rng, np.float64, (a, s1))
    t2 = B.randn_itensor(rng, np.float64, (b, s2, a))
    t3 = B.randn_itensor(rng, np.float64, (s3, b))
    tn = B.ITensorNetwork({1: t1, 2: t2, 3: t3})
    cp = B.canonical_arrays(B.normnetwork(tn))
    assert cp.tensors[0].shape == (2, 2) and np.array_equal(cp.tensors[0], t1.data.T)
    # vertex 2: neighbours in leg order (b -> 3, a -> 1)
    assert [cp.ga.vertices[cp.ga.dst[e]] for e in range(cp.ga.row_ptr[1], cp.ga.row_ptr[2])] == [3, 1]
    assert np.array_equal(cp.tensors[1], np.transpose(t2.data, (1, 0, 2)))


def test_grid_graph_arrays_matches_named_grid():
    """The vectorised lattice builder (millions of vertices) describes the same graph as named_grid + graph_arrays."""
    for dims, periodic in [((4, 4), False), ((4, 4), True), ((3, 2), True), ((3, 3, 3), True), ((5,), True), ((2, 2), True), ((1, 4), False)]:
        ga = graphs.grid_graph_arrays(dims, periodic)
        gb = graphs.graph_arrays(graphs.named_grid(dims, periodic))
        assert ga.nv == gb.nv and ga.ne == gb.ne
        assert set(zip(ga.src.tolist(), ga.dst.tolist())) == set(zip(gb.src, gb.dst))  # same vertex numbering
        assert np.all(ga.src[ga.rev] == ga.dst) and np.all(ga.rev[ga.rev] == np.arange(ga.ne))
        assert np.all(ga.slot == np.arange(ga.ne) - ga.row_ptr[ga.src]) and np.all(np.diff(ga.src) >= 0)
    big = graphs.grid_graph_arrays((1024, 1024), True)
    assert big.nv == 1 << 20 and big.ne == 1 << 22 and np.all(np.diff(big.row_ptr) == 4)


def test_synthetic_ising_workload_is_the_generator_network(oracle):
    """bench.py --workload ising builds its factors in bulk; on a small lattice they contract to the same partition
    function as the `ising_network` generator's network, and the work model of the roofline is the documented one."""
    from itnn_b200 import generators, problems
    from itnn_b200.tensornetwork import Index, canonical_arrays

    for dims, periodic in [((4, 4), True), ((4, 3), False)]:
        q = problems.synthetic_ising(dims, 0.3, periodic)
        tensors, msgs = problems.unpacked(q)
        z1 = oracle.contract_all_sequential(oracle.make_problem(q.ga, tensors, "single"))
        g = graphs.named_grid(dims, periodic)
        ld = {frozenset((e.src, e.dst)): Index(2) for e in g.edges()}
        cp = canonical_arrays(generators.ising_network(lambda e: ld[frozenset((e.src, e.dst))], 0.3, g))
        z2 = oracle.contract_all_sequential(oracle.make_problem(cp.ga, cp.tensors, "single"))
        assert np.isclose(z1, z2, rtol=1e-12)
        assert all(abs(m.sum() - 1) < 1e-12 and (m > 0).all() for m in msgs)
    q = problems.synthetic_ising((8, 8))
    assert q.mode == "single" and q.bytes_per_sweep() == 64 * 16 * 8 + 3 * 256 * 2 * 8   # factors + 3 x messages
    assert q.flops_per_sweep() == 256 * 2 * (16 + 8 + 4)                                # absorb 3 legs: 16 + 8 + 4 MACs


def test_pack_operators_stacked_fast_path_equals_the_per_operator_loop():
    """device.pack_operators: a batch of same-shaped operators is packed in one stacked transpose; ragged batches, lists and
    dtype promotion go through the per-operator loop -- same bytes either way, and a complex operator for a Float64
    context is still refused."""
    import pytest

    from itnn_b200.device import cast_to, pack_operators

    rng = np.random.default_rng(0)
    for dt_in, dt_out in ((np.float64, np.float64), (np.float64, np.complex128), (np.complex128, np.complex128)):
        for shape in ((2, 2, 2, 2), (2, 3, 2, 3), (3, 3)):
            ops = [(rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if dt_in == np.complex128 else 0)).astype(dt_in)
                   for _ in range(7)]
            want = np.concatenate([cast_to(o, dt_out, "operator").ravel(order="F") for o in ops])
            got = pack_operators(ops, dt_out)
            assert got.dtype == np.dtype(dt_out) and got.flags.c_contiguous and np.array_equal(got, want)
    ragged = [rng.standard_normal((2, 2, 2, 2)), rng.standard_normal((3, 3, 3, 3))]
    assert np.array_equal(pack_operators(ragged, np.float64), np.concatenate([o.ravel(order="F") for o in ragged]))
    assert pack_operators([], np.float64).size == 0
    assert np.array_equal(pack_operators([[[1.0, 2.0], [3.0, 4.0]]], np.float64), [1.0, 3.0, 2.0, 4.0])  # nested lists
    with pytest.raises(TypeError):
        pack_operators([np.ones((2, 2), np.complex128) * 1j] * 3, np.float64)
