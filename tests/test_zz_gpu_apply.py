"""GPU parity of the gate-application kernels through the C ABI (bpx_apply_two_site_gates / bpx_apply_one_site_gates,
SURVEY.md §8 f4) against oracle/apply_oracle.py (which restates src/apply/apply_operators.jl:213-283), in one context
with the BP sweeps: sweep -> gates on a layer of disjoint edges -> sweep again (the update kernels' private tensor
images must follow the new tensors)."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn
from itnn_b200 import graphs, problems
from oracle import apply_oracle as A

pytestmark = pytest.mark.gpu
DTYPES = [np.float64, np.complex128]


def oracle_state(p, tensors):
    link = lambda v, w: ("l", min(v, w), max(v, w))
    state = {}
    for v in range(p.ga.nv):
        nb = [p.ga.dst[e] for e in range(p.ga.row_ptr[v], p.ga.row_ptr[v + 1])]
        state[v] = (np.asarray(tensors[v]), (("s", v),) + tuple(link(v, w) for w in nb))
    return state


def matching(ga, rng):
    """A maximal set of vertex-disjoint directed edges, in random order and orientation."""
    used, out = set(), []
    for e in rng.permutation(ga.ne):
        s, d = ga.src[e], ga.dst[e]
        if s not in used and d not in used:
            used |= {s, d}
            out.append(int(e))
    return out


def device_tensors(ctx, p):
    shapes = [(p.d,) + tuple(p.link_dim[e] for e in range(p.ga.row_ptr[v], p.ga.row_ptr[v + 1])) for v in range(p.ga.nv)]
    return [ctx.get_site_tensor(v).reshape(shapes[v], order="F") for v in range(p.ga.nv)]


def bond_invariant(state, v1, v2):
    t = A.contract(state[v1], state[v2])
    return A.permute(t, sorted(t[1], key=repr))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("lattice,chi,max_rank,normalize", [((4, 4), 4, 0, False), ((4, 4), 4, 2, True), ((3, 5), 3, 3, True),
                                                            ((4, 4), 8, 8, False)])
def test_two_site_layer_matches_oracle(oracle, dtype, lattice, chi, max_rank, normalize):
    rng = np.random.default_rng(chi * 7 + max_rank)
    p = problems.synthetic_peps(graphs.named_grid(lattice), chi, 2, dtype, init="positive")
    ga = p.ga
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(6, 0.0, True)                       # messages of a BP run in progress: full, nearly Hermitian PSD
        msgs = ctx.get_messages()
        edges = matching(ga, rng)
        assert len(edges) >= 4
        ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
        svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=normalize)
        new_msgs = ctx.get_messages()
        new_tensors = device_tensors(ctx, p)
        launches = ctx.counters()["launches"]
        # the BP kernels must see the new tensors (their private pre-swizzled images are rebuilt)
        res, _ = ctx.sweep(1, 0.0, True)
        swept = ctx.get_messages()

    state = oracle_state(p, p.tensors)
    env = {(ga.src[e], ga.dst[e]): msgs[e] for e in range(ga.ne)}
    k_want = max_rank or chi
    got_state = oracle_state(p, new_tensors)
    touched_edges = set()
    for e, op, sv in zip(edges, ops, svs):
        v1, v2, r = ga.src[e], ga.dst[e], ga.rev[e]
        names = (("s", v1), ("s", v2))
        want_state, want_env = A.apply_operator((op, names, names), state, env, trunc=k_want, normalize=normalize)
        s_want = np.diag(want_env[(v1, v2)]).real
        k = len(s_want)
        assert np.allclose(sv[:k], s_want, rtol=1e-9, atol=1e-13) and np.all(sv[k:] == 0)
        for ee in (e, r):
            assert np.allclose(new_msgs[ee][:k, :k], np.diag(s_want), rtol=1e-9, atol=1e-13)
            assert np.all(new_msgs[ee][k:, :] == 0) and np.all(new_msgs[ee][:, k:] == 0)
        touched_edges |= {e, r}
        # zero-pad the oracle's (possibly smaller) bond for the comparison of the gauge-invariant pair product
        got, want = bond_invariant(got_state, v1, v2), bond_invariant(want_state, v1, v2)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    # everything else is untouched, bit for bit
    in_gate = {ga.src[e] for e in edges} | {ga.dst[e] for e in edges}
    for v in range(ga.nv):
        if v not in in_gate:
            assert np.array_equal(new_tensors[v], p.tensors[v])
    for e in range(ga.ne):
        if e not in touched_edges:
            assert np.array_equal(new_msgs[e], msgs[e])
    assert launches >= 7  # 6 sweeps + the gate layer (three launches on the Gram path)
    # one more sweep from the device's own new tensors and messages
    op_ = oracle.make_problem(ga, new_tensors, "norm")
    want_swept = oracle.sweep_jacobi(op_, new_msgs)
    err = max(np.abs(g - w).max() / max(np.abs(w).max(), 1e-300) for g, w in zip(swept, want_swept))
    assert err < 1e-10
    assert abs(res - oracle.iterate_diff(want_swept, new_msgs)) < 1e-11


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("normalize", [False, True])
def test_one_site_layer_matches_oracle(dtype, normalize):
    rng = np.random.default_rng(3)
    p = problems.synthetic_peps(graphs.named_grid((3, 4)), 3, 2, dtype, init="positive")
    ga = p.ga
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(4, 0.0, True)
        msgs = ctx.get_messages()
        vertices = [int(v) for v in rng.permutation(ga.nv)[: ga.nv - 2]]
        ops = [randn(rng, dtype, (2, 2)) for _ in vertices]
        ctx.apply_one_site_gates(vertices, ops, normalize=normalize)
        new_tensors = device_tensors(ctx, p)
        assert all(np.array_equal(a, b) for a, b in zip(ctx.get_messages(), msgs))
    state = oracle_state(p, p.tensors)
    env = {(ga.src[e], ga.dst[e]): msgs[e] for e in range(ga.ne)}
    for v, op in zip(vertices, ops):
        names = (("s", v),)
        want_state, _ = A.apply_operator((op, names, names), state, env, normalize=normalize)
        want = A.permute(want_state[v], state[v][1])
        assert np.abs(new_tensors[v] - want).max() <= 1e-11 * np.abs(want).max()
    for v in set(range(ga.nv)) - set(vertices):
        assert np.array_equal(new_tensors[v], p.tensors[v])


def test_apply_argument_errors():
    p = problems.synthetic_peps(graphs.named_grid((3, 3)), 2, 2, np.float64)
    ga = p.ga
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        e0 = 0
        e1 = next(e for e in range(ga.ne) if ga.src[e] == ga.dst[e0] and e != ga.rev[e0])  # shares a vertex with e0
        op = np.eye(4).reshape(2, 2, 2, 2)
        with pytest.raises(B.BPXError, match="vertex-disjoint") as ei:
            ctx.apply_two_site_gates([e0, e1], [op, op])
        assert ei.value.status == -1
        with pytest.raises(B.BPXError, match="out of range"):
            ctx.apply_two_site_gates([ga.ne], [op])
        with pytest.raises(B.BPXError, match="twice"):
            ctx.apply_one_site_gates([1, 1], [np.eye(2), np.eye(2)])
        assert ctx.apply_two_site_gates([], []) == []  # an empty layer is a no-op
        # the identity gate with the full rank kept leaves the state invariant (up to the bond gauge)
        before = [ctx.get_site_tensor(v) for v in range(ga.nv)]
        ctx.apply_two_site_gates([e0], [op], max_rank=0)
        v1, v2 = ga.src[e0], ga.dst[e0]
        shapes = lambda v: (2,) + tuple(p.link_dim[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1]))
        old = oracle_state(p, [b.reshape(shapes(v), order="F") for v, b in enumerate(before)])
        new = oracle_state(p, [ctx.get_site_tensor(v).reshape(shapes(v), order="F") for v in range(ga.nv)])
        assert np.allclose(bond_invariant(new, v1, v2), bond_invariant(old, v1, v2), rtol=1e-10, atol=1e-13)
    single = problems.synthetic_ising((4, 4))
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, single)
        with pytest.raises(B.BPXError, match="NORM mode"):
            ctx.apply_one_site_gates([0], [np.eye(1)])
