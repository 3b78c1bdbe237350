"""`apply_operator` / `apply_operators` of the host mirror (itensornetworksnext.jl_b200/apply.py) against the
reference's own known-answer tests (test/test_apply_operator.jl:62-133) and the apply oracle.

  * `-m "not gpu"`: the mirror's lowering (names -> canonical layout, bond padding / slicing, layer batching) with the
    device code executed on the host through tests/native_ctx.HostHarnessContext (csrc/bpx_apply.cuh compiled for the
    host; test infrastructure, never a product path).
  * `-m gpu`: the same checks through libbpx.so -- BP messages AND gate application on the B200.
"""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn
from itnn_b200 import apply as apply_mod
from itnn_b200 import graphs
from native_ctx import HostHarnessContext
from oracle import apply_oracle as A
from test_apply_device_code import hostlib  # noqa: F401  (fixture: builds tests/native/apply_host.cu)
from test_apply_oracle import (DTYPES, bp_environment_oracle, check_known_answers, env_from_cache, random_state,
                               randn_operator, scramble_gauges, site_name, to_network)


# ---- oracle calling convention <-> mirror ----------------------------------------------------------------------
def to_cache(net, env):
    out = {}
    for (w, v), m in env.items():
        l = net.linkname(B.NamedEdge(w, v))
        c = m.shape[0]
        out[B.NamedEdge(w, v)] = B.ITensor(np.asarray(m), (B.Index(c, ("bra", l)), B.Index(c, l)))
    return B.MessageCache(out)


def from_network(net):
    return {v: (net[v].data, net[v].dimnames()) for v in net.vertices()}


def from_cache(net, cache):
    env = {}
    for e, m in cache.items():
        ket = net.linkname(e)
        bra = next(n for n in m.dimnames() if n != ket)
        env[(e.src, e.dst)] = m.array(bra, ket)
    return env


def mirror_apply_operator(op, state, env, trunc=None, normalize=False):
    net = to_network(state)
    new_net, new_cache = B.apply_operator(B.Operator(*op), net, to_cache(net, env), trunc=trunc, normalize=normalize)
    return from_network(new_net), from_cache(new_net, new_cache)


def mirror_apply_operators(ops, state, env, **kwargs):
    net = to_network(state)
    new_net, new_cache = B.apply_operators([B.Operator(*op) for op in ops], net, to_cache(net, env), **kwargs)
    return from_network(new_net), from_cache(new_net, new_cache)


@pytest.fixture(params=[None, 2, 64], ids=["v1", "v2-rb2", "v2-rb64"])
def host_ctx(monkeypatch, hostlib, request):  # noqa: F811
    made = []

    def factory(device=0):
        c = HostHarnessContext(hostlib, device, v2_block_rows=request.param)
        made.append(c)
        return c

    monkeypatch.setattr(apply_mod, "BPXContext", factory)
    return made


# ---- shared checks ----------------------------------------------------------------------------------------------
def bond_invariant(state, v1, v2):
    t = A.contract(state[v1], state[v2])
    return A.permute(t, sorted(t[1], key=repr))


def check_against_oracle_on_a_grid(dtype):
    """3 x 3 grid PEPS with generic (full, complex Hermitian PSD) environments: one- and two-site gates, with and
    without truncation / normalisation, and a layer of disjoint gates, against oracle/apply_oracle.py."""
    rng = np.random.default_rng(5)
    g = graphs.named_grid((3, 3))
    net, _, _ = B.random_state(dtype, g, d=2, chi=3, rng=rng)
    state = from_network(net)
    env = {}
    for e in g.all_edges():
        f = randn(rng, dtype, (3, 3)) + 1.5 * np.eye(3)
        m = f.conj().T @ f
        env[(e.src, e.dst)] = (m / np.trace(m).real).astype(dtype)
    sites = {v: net.sitenames(v)[0] for v in g.vertices()}

    def gate(*vs):
        names = tuple(sites[v] for v in vs)
        return randn(rng, dtype, (2,) * (2 * len(vs))), names, names

    for vs, kw in [(((2, 2), (2, 3)), dict(trunc=3)), (((1, 1), (2, 1)), dict(trunc=2, normalize=True)),
                   (((2, 2), (1, 2)), dict()), (((2, 2),), dict(normalize=True)), (((3, 3),), dict())]:
        op = gate(*vs)
        want_state, want_env = A.apply_operator(op, state, env, **kw)
        got_state, got_env = mirror_apply_operator(op, state, env, **kw)
        for v in state:
            assert got_state[v][0].shape == A.permute(want_state[v], got_state[v][1]).shape, (vs, v)
        if len(vs) == 2:
            got, want = bond_invariant(got_state, *vs), bond_invariant(want_state, *vs)
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
            for key in (vs, vs[::-1]):
                assert np.allclose(got_env[key], want_env[key], rtol=1e-9, atol=1e-12)
        else:
            want = A.permute(want_state[vs[0]], got_state[vs[0]][1])
            assert np.abs(got_state[vs[0]][0] - want).max() <= 1e-11 * np.abs(want).max()
        untouched = [v for v in state if v not in vs]
        assert all(np.array_equal(got_state[v][0], state[v][0]) for v in untouched)
        assert all(np.array_equal(got_env[k], env[k]) for k in env if set(k) != set(vs))

    # one layer of disjoint gates followed by an overlapping one
    ops = [gate((1, 1), (1, 2)), gate((2, 1), (2, 2)), gate((3, 2), (3, 3)), gate((1, 2), (2, 2))]
    want_state, want_env = A.apply_operators(ops, state, env, trunc=3)
    got_state, got_env = mirror_apply_operators(ops, state, env, trunc=3)
    full_got, full_want = A.prod(got_state), A.prod(want_state)
    assert np.abs(A.permute(full_got, full_want[1]) - full_want[0]).max() <= 1e-9 * np.abs(full_want[0]).max()
    for key in want_env:
        assert np.allclose(got_env[key], want_env[key], rtol=1e-9, atol=1e-12)


def check_errors():
    rng = np.random.default_rng(1)
    g = graphs.named_path_graph(4)
    state = random_state(rng, np.float64, g)
    net = to_network(state)
    env = to_cache(net, {(e.src, e.dst): np.eye(state[e.src][0].shape[state[e.src][1].index(A.linkname(state, e.src, e.dst))])
                         for e in g.all_edges()})
    with pytest.raises(B.ArgumentError, match="shares no indices"):
        B.apply_operator(B.Operator(np.eye(3), [("s", 99)], [("s", 99)]), net, env)
    three = tuple(site_name(v) for v in (1, 2, 3))
    with pytest.raises(B.ArgumentError, match="3-site gate decomposition not implemented"):
        B.apply_operator(B.Operator(np.zeros((3,) * 6), three, three), net, env)
    far = (site_name(1), site_name(3))
    with pytest.raises(B.ArgumentError, match="share no link"):
        B.apply_operator(B.Operator(np.zeros((3,) * 4), far, far), net, env)
    with pytest.raises(B.ArgumentError):
        B.apply_operator(B.Operator(np.eye(3), [site_name(1)], [site_name(1)]), net, env, alg=B.BPApplyGate(), trunc=2)
    out_net, out_env = B.apply_operators([], net, env)  # apply_operators.jl:55: copies
    assert out_net is not net and all(np.array_equal(out_net[v].data, net[v].data) for v in net.vertices())


def check_expect_two_site(oracle, dtype):
    """`expect_two_site` of the mirror (names -> canonical layout -> bpx_edge_expect) against the oracle on a 3 x 3 PEPS
    whose site legs are NOT the first axis of the stored tensors."""
    rng = np.random.default_rng(12)
    g = graphs.named_grid((3, 3))
    net, _, sites = B.random_state(dtype, g, d=2, chi=2, rng=rng)
    # store every tensor with its site leg last: the lowering must permute by name
    net = B.ITensorNetwork({v: B.ITensor(np.moveaxis(net[v].data, 0, -1), net[v].inds[1:] + net[v].inds[:1]) for v in net.vertices()})
    nn = B.normnetwork(net)
    cp = B.canonical_arrays(nn)
    env_arrays = []
    for e in range(cp.ga.ne):
        f = randn(rng, dtype, (2, 2)) + 1.5 * np.eye(2)
        env_arrays.append((f.conj().T @ f).astype(dtype))
    cache = B.MessageCache({cp.ga.named_edge(e): B.ITensor(env_arrays[e], (B.Index(2, ("bra", cp.ket_names[e])), B.Index(2, cp.ket_names[e])))
                            for e in range(cp.ga.ne)})
    pairs = [((1, 1), (2, 1)), ((2, 2), (2, 3)), ((2, 3), (2, 2)), ((3, 3), (3, 2))]
    ops = []
    for v, w in pairs:
        names = (sites[v].name, sites[w].name)
        ops.append(B.Operator(randn(rng, dtype, (2, 2, 2, 2)), names, names))
    got = B.expect_two_site(ops, net, cache)
    p = oracle.make_problem(cp.ga, cp.tensors, "norm")
    for (v, w), op, val in zip(pairs, ops, got):
        # the lowering orders the two vertices as they appear in `vertices(state)`
        first, second = (v, w) if cp.ga.vindex[v] < cp.ga.vindex[w] else (w, v)
        o = op.data if (first, second) == (v, w) else np.transpose(op.data, (1, 0, 3, 2))
        num, den = oracle.two_site_expect(p, env_arrays, cp.ga.edge_id(B.NamedEdge(first, second)), o)
        assert np.isclose(val, num / den, rtol=1e-10, atol=1e-13)
    with pytest.raises(B.ArgumentError, match="neighbouring"):
        names = (sites[(1, 1)].name, sites[(3, 3)].name)
        B.expect_two_site([B.Operator(np.zeros((2, 2, 2, 2)), names, names)], net, cache)


# ---- CPU: lowering + device code on the host -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_mirror_known_answers_host_harness(oracle, host_ctx, dtype):
    check_known_answers(lambda state, g: bp_environment_oracle(oracle, state, g), dtype,
                        apply_operator=mirror_apply_operator, apply_operators=mirror_apply_operators)
    assert host_ctx, "the mirror did not go through its context"


@pytest.mark.parametrize("dtype", DTYPES)
def test_mirror_matches_oracle_on_a_grid_host_harness(host_ctx, dtype):
    check_against_oracle_on_a_grid(dtype)
    # the layer of three disjoint gates went to the device in one call, the overlapping gate in the next
    assert [c.calls for c in host_ctx[-2:]] == [[("two", 3)], [("two", 1)]]


def test_mirror_errors_host_harness(host_ctx):
    check_errors()


@pytest.mark.parametrize("dtype", DTYPES)
def test_mirror_expect_two_site_host_harness(oracle, host_ctx, dtype):
    check_expect_two_site(oracle, dtype)


# ---- GPU: BP messages and gate application on the B200 ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_mirror_known_answers_gpu(dtype):
    def env_of(state, g):
        nn = B.normnetwork(to_network(state))
        env0 = B.message_environment(B.ones_message, nn)
        cache = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=100, tol=1e-13))
        return env_from_cache(nn, cache)

    check_known_answers(env_of, dtype, apply_operator=mirror_apply_operator, apply_operators=mirror_apply_operators)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_mirror_matches_oracle_on_a_grid_gpu(dtype):
    check_against_oracle_on_a_grid(dtype)


@pytest.mark.gpu
def test_mirror_errors_gpu():
    check_errors()


# ---- dtype promotion: a complex gate on a real state must not lose its imaginary part -------------
def test_complex_gate_on_a_float64_state_promotes_host_harness(host_ctx):
    rng = np.random.default_rng(11)
    g = graphs.named_path_graph(3)
    net, _, sites = B.random_state(np.float64, g, d=2, chi=2, rng=rng)
    cache = B.MessageCache(B.message_environment(B.ones_message, B.normnetwork(net)))
    names = (sites[1].name, sites[2].name)
    h = rng.standard_normal((4, 4))
    h = h + h.T
    w, v = np.linalg.eigh(h)
    u = (v * np.exp(-0.3j * w)) @ v.conj().T                     # exp(-i dt H): a genuinely complex two-site gate
    op = B.Operator(u.reshape(2, 2, 2, 2), names, names)
    got_net, got_env = B.apply_operator(op, net, cache)
    assert all(np.iscomplexobj(got_net[x].data) for x in (1, 2))
    # same gate on the explicitly promoted state
    cnet = B.ITensorNetwork({x: B.ITensor(net[x].data.astype(np.complex128), net[x].inds) for x in net.vertices()})
    want_net, want_env = B.apply_operator(op, cnet, cache)
    for x in net.vertices():
        assert np.allclose(got_net[x].data, want_net[x].data, rtol=0, atol=1e-13)
    assert np.abs(np.asarray(got_net[1].data).imag).max() > 1e-3    # the imaginary part is really there
    # one-site gate, same story
    op1 = B.Operator(np.array([[1.0, 0.0], [0.0, 1.0j]]), (sites[1].name,), (sites[1].name,))
    n1, _ = B.apply_operator(op1, net, cache)
    assert np.iscomplexobj(n1[1].data) and np.abs(n1[1].data.imag).max() > 0


def test_context_refuses_to_drop_an_imaginary_part():
    from itnn_b200.device import cast_to

    assert cast_to(np.array([1.0 + 0.0j, 2.0]), np.float64, "operator").dtype == np.float64   # harmless: imaginary part is zero
    with pytest.raises(TypeError, match="imaginary part"):
        cast_to(np.array([1.0 + 1.0j]), np.float64, "operator")
    assert cast_to(np.array([1.0, 2.0]), np.complex128, "message").dtype == np.complex128
