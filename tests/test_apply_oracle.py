"""The reference's apply_operator known-answer tests (test/test_apply_operator.jl:62-133) on the CPU oracle: they pin the
BP messages of a NormNetwork -- a truncated two-site gate on an open chain reproduces the globally optimal truncated SVD
only if the messages are the true environments with the [bra, ket] orientation right.

`gpu`-marked variants take the messages from the CUDA path instead of the oracle's BP."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn
from itnn_b200 import graphs
from oracle import apply_oracle as A

DTYPES = [np.float64, np.complex128]
D_SITE = 3  # spin one (test/test_apply_operator.jl:16)


def site_name(v):
    return ("s", v)


def randn_operator(rng, dtype, vs):
    """test/test_apply_operator.jl:19-27: axes (out..., in...), i.i.d. normal."""
    names = tuple(site_name(v) for v in vs)
    return randn(rng, dtype, (D_SITE,) * (2 * len(vs))), names, names


def random_state(rng, dtype, g, nlayers=2, trunc=4):
    """test/test_apply_operator.jl:29-45: random product state, trivial links, then `nlayers` of random two-site gates
    on every edge, truncated to rank `trunc`, environments updated by the gates only."""
    link = {frozenset((e.src, e.dst)): ("l", e.src, e.dst) for e in g.edges()}
    state = {}
    for v in g.vertices():
        names = (site_name(v),) + tuple(link[frozenset((v, w))] for w in g.neighbors(v))
        state[v] = (randn(rng, dtype, (D_SITE,) + (1,) * g.degree(v)), names)
    env = {(e.src, e.dst): np.ones((1, 1), dtype=dtype) for e in g.all_edges()}
    for _ in range(nlayers):
        for e in g.edges():
            state, env = A.apply_operator(randn_operator(rng, dtype, (e.src, e.dst)), state, env, trunc=trunc)
    return state


def scramble_gauges(rng, dtype, state, g):
    """Insert G G^-1 (random invertible G) on every link: the same physical state in a generic gauge, so that the BP
    messages are full (complex Hermitian) matrices instead of the diagonal ones a simple-update history leaves behind."""
    state = dict(state)
    for e in g.edges():
        l = A.linkname(state, e.src, e.dst)
        n = state[e.src][0].shape[state[e.src][1].index(l)]
        gm = randn(rng, dtype, (n, n)) + 2.0 * np.eye(n)
        tmp = A.fresh("t")
        state[e.src] = A.rename(A.contract(state[e.src], (gm, (l, tmp))), {tmp: l})
        state[e.dst] = A.rename(A.contract(state[e.dst], (np.linalg.inv(gm), (tmp, l))), {tmp: l})
    return state


def to_network(state):
    """oracle state -> the host mirror's ITensorNetwork (names and dims only change container)."""
    return B.ITensorNetwork({v: B.ITensor(np.asfortranarray(x), [B.Index(d, n) for d, n in zip(x.shape, names)])
                             for v, (x, names) in state.items()})


def bp_environment_oracle(oracle, state, g):
    """BP on NormNetwork(state) from all-ones messages, maxiter 100, tol 1e-13 (test/test_apply_operator.jl:70-74),
    the reference's sequential schedule."""
    nn = B.normnetwork(to_network(state))
    cp = B.canonical_arrays(nn)
    p = oracle.make_problem(cp.ga, cp.tensors, "norm")
    msgs = [np.ones((c, c), dtype=cp.dtype) for c in cp.link_dim]
    seq = [cp.ga.edge_id(e) for e in graphs.forest_cover_edge_sequence(nn.graph)]
    out, it, delta = oracle.beliefpropagation(p, msgs, maxiter=100, tol=1e-13, schedule="sequential", edge_seq=seq)
    assert delta < 1e-13
    return {(cp.ga.named_edge(e).src, cp.ga.named_edge(e).dst): out[e] for e in range(cp.ga.ne)}


def env_from_cache(nn, cache):
    """A MessageCache of ITensors (what `beliefpropagation` returns) -> {(w, v): M[bra, ket]}."""
    env = {}
    for e in nn.graph.all_edges():
        ket = nn.ket.linkname(e)
        env[(e.src, e.dst)] = cache[e].array(nn.braname(ket), ket)
    return env


def full_tensor(state, order):
    return A.permute(A.prod(state), order)


def check_known_answers(env_of, dtype, apply_operator=A.apply_operator, apply_operators=A.apply_operators):
    """`apply_operator(s)`: the implementation under test, in the oracle's calling convention (default: the oracle itself)."""
    rtol = np.finfo(np.float64).eps ** (1 / 3)
    n = 4
    sites = [site_name(v) for v in range(1, n + 1)]

    def close(a, b):
        return np.linalg.norm((a - b).ravel()) <= rtol * max(np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel()))

    # "untruncated gates are exact (gauge-invariant)": cycle graph, one- and two-site gate (:62-85)
    rng = np.random.default_rng(123)
    g = graphs.named_cycle_graph(n)
    state = random_state(rng, dtype, g)
    env = env_of(state, g)
    for vs in ((2,), (2, 3)):
        gate = randn_operator(rng, dtype, vs)
        gated, _ = apply_operator(gate, state, env)
        want = A.permute(A.apply_op(gate, A.prod(state)), sites)
        assert close(full_tensor(gated, sites), want)

    # "truncated 2-site gate matches global optimal SVD (rank k)": open chain (:87-110)
    for k, scrambled in ((1, False), (2, False), (3, False), (1, True), (2, True)):
        rng = np.random.default_rng(123)
        g = graphs.named_path_graph(n)
        state = random_state(rng, dtype, g)
        if scrambled:  # beyond the reference's test: a generic gauge, where the messages are full matrices
            state = scramble_gauges(rng, dtype, state, g)
        env = env_of(state, g)
        gate = randn_operator(rng, dtype, (2, 3))
        full = A.permute(A.apply_op(gate, A.prod(state)), sites)
        u, s, vh = np.linalg.svd(full.reshape(D_SITE ** 2, D_SITE ** 2), full_matrices=False)
        want = ((u[:, :k] * s[:k]) @ vh[:k]).reshape(full.shape)
        gated, new_env = apply_operator(gate, state, env, trunc=k)
        assert close(full_tensor(gated, sites), want)
        assert new_env[(2, 3)].shape == (k, k) and np.allclose(new_env[(2, 3)], new_env[(3, 2)])
        # all-ones messages are NOT the environments: the same truncation with them is not optimal (the test has teeth)
        if k < 3:
            ones = {key: np.ones_like(m) for key, m in env.items()}
            worse, _ = apply_operator(gate, state, ones, trunc=k)
            assert not close(full_tensor(worse, sites), want)

    # "apply_operators applies a sequence" (:112-133)
    rng = np.random.default_rng(123)
    g = graphs.named_cycle_graph(n)
    state = random_state(rng, dtype, g)
    env = env_of(state, g)
    g1, g2 = randn_operator(rng, dtype, (2, 3)), randn_operator(rng, dtype, (3, 4))
    gated, _ = apply_operators([g1, g2], state, env)
    want = A.permute(A.apply_op(g2, A.apply_op(g1, A.prod(state))), sites)
    assert close(full_tensor(gated, sites), want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_apply_operator_known_answers_with_oracle_bp(oracle, dtype):
    check_known_answers(lambda state, g: bp_environment_oracle(oracle, state, g), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_message_orientation_matters_for_complex_states(oracle, dtype):
    """Transposed messages ([ket, bra]) are the environments of the conjugate state: for complex states the truncated gate
    stops being optimal, for real states nothing changes -- so the known answer above pins the [bra, ket] layout."""
    rng = np.random.default_rng(123)
    g = graphs.named_path_graph(4)
    state = scramble_gauges(rng, dtype, random_state(rng, dtype, g), g)
    env = bp_environment_oracle(oracle, state, g)
    assert max(np.abs(m - np.diag(np.diag(m))).max() for m in env.values()) > 1e-3  # generic gauge: not diagonal
    sites = [site_name(v) for v in range(1, 5)]
    gate = randn_operator(rng, dtype, (2, 3))
    good, _ = A.apply_operator(gate, state, env, trunc=2)
    flipped, _ = A.apply_operator(gate, state, {k: m.T.copy() for k, m in env.items()}, trunc=2)
    diff = np.linalg.norm((full_tensor(good, sites) - full_tensor(flipped, sites)).ravel()) / np.linalg.norm(full_tensor(good, sites).ravel())
    if np.dtype(dtype).kind == "c":
        assert diff > 1e-6
    else:
        assert diff < 1e-10


def test_one_site_gate_normalisation_and_errors(oracle):
    rng = np.random.default_rng(7)
    g = graphs.named_path_graph(4)
    state = random_state(rng, np.complex128, g)
    env = bp_environment_oracle(oracle, state, g)
    gate = randn_operator(rng, np.complex128, (2,))
    plain, _ = A.apply_operator(gate, state, env)
    normed, _ = A.apply_operator(gate, state, env, normalize=True)
    ratio = plain[2][0] / normed[2][0]
    assert np.allclose(ratio, ratio.flat[0]) and abs(ratio.flat[0].imag) < 1e-12
    # with the true environments the gauged norm is the norm of the whole gated state (up to the BP normalisation
    # of the messages, which is the same for every gate): two different gates give the same ratio of norms
    gate2 = randn_operator(rng, np.complex128, (2,))
    plain2, _ = A.apply_operator(gate2, state, env)
    normed2, _ = A.apply_operator(gate2, state, env, normalize=True)
    r1 = np.linalg.norm(A.prod(plain)[0].ravel()) / ratio.flat[0].real
    r2 = np.linalg.norm(A.prod(plain2)[0].ravel()) / (plain2[2][0] / normed2[2][0]).flat[0].real
    assert np.isclose(r1, r2, rtol=1e-9)
    with pytest.raises(ValueError):
        A.apply_operator((np.eye(3), (("s", 99),), (("s", 99),)), state, env)


def test_env_from_cache_reads_bra_ket_by_name(oracle):
    """The glue the GPU variant uses: MessageCache of ITensors -> [bra, ket] arrays, whatever the stored axis order."""
    rng = np.random.default_rng(3)
    g = graphs.named_path_graph(3)
    state = random_state(rng, np.complex128, g)
    nn = B.normnetwork(to_network(state))
    want = bp_environment_oracle(oracle, state, g)
    msgs = {}
    for i, e in enumerate(nn.graph.all_edges()):
        ket = nn.ket.linkname(e)
        m = want[(e.src, e.dst)]
        kb = (B.Index(m.shape[0], nn.braname(ket)), B.Index(m.shape[1], ket))
        msgs[e] = B.ITensor(m, kb) if i % 2 == 0 else B.ITensor(m.T.copy(), kb[::-1])
    got = env_from_cache(nn, B.MessageCache(msgs))
    assert all(np.array_equal(got[k], want[k]) for k in want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_apply_operator_known_answers_with_gpu_bp(oracle, dtype):
    """Same known answers with the messages computed by the CUDA path through the reference-shaped API."""

    def env_of(state, g):
        nn = B.normnetwork(to_network(state))
        env0 = B.message_environment(B.ones_message, nn)
        cache = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=100, tol=1e-13))
        return env_from_cache(nn, cache)

    check_known_answers(env_of, dtype)
