"""The gate-application device code (csrc/bpx_apply.cuh) checked WITHOUT a GPU: tests/native/apply_host.cu
compiles the same `__host__ __device__` routines for the host (a CUDA block collapses to one sequential lane), and
every stage -- one-sided Jacobi, Householder QR, gauges from messages, the whole one- and two-site gate -- is compared
with numpy / the apply_operator oracle (oracle/apply_oracle.py, which restates src/apply/apply_operators.jl:213-283).
The CUDA launch of the same code is covered by the `gpu` tests in test_zz_gpu_apply.py."""
import ctypes

import numpy as np
import pytest

from helpers import randn
from native_ctx import build_hostlib
from oracle import apply_oracle as A

DTYPES = [np.float64, np.complex128]
P = ctypes.c_void_p


@pytest.fixture(scope="module")
def hostlib():
    return build_hostlib()


def ptr(a):
    return a.ctypes.data_as(P)


def code(dtype):
    return 1 if np.dtype(dtype).kind == "c" else 0


def fcopy(a):
    return np.array(a, order="F", copy=True)


# ---------------------------------------------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(6, 6), (9, 4), (4, 9), (12, 7), (1, 3), (5, 1), (32, 32), (64, 64), (64, 32), (32, 64)])
def test_jacobi_svd(hostlib, dtype, shape):
    rng = np.random.default_rng(11)
    m, n = shape
    a = randn(rng, dtype, (m, n))
    if m >= 6 and n >= 4:
        a[:, 1] = 0.5 * a[:, 0]  # a rank-deficient pair
    b, v = fcopy(a), np.zeros((n, n), dtype=dtype, order="F")
    hostlib.apply_host_jacobi(code(dtype), ptr(b), m, n, ptr(v))
    assert np.allclose(v.conj().T @ v, np.eye(n), atol=1e-13)          # V unitary
    assert np.allclose(a @ v, b, atol=1e-12 * np.abs(a).max())          # A V = B
    gram = b.conj().T @ b
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-13 * np.abs(gram).max()              # orthogonal columns
    got = np.sort(np.sqrt(np.diag(gram).real))[::-1][:min(m, n)]
    want = np.linalg.svd(a, compute_uv=False)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13 * want.max())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(20, 6), (6, 6), (4, 6), (1, 4), (2, 5), (64, 8)])
def test_householder_qr(hostlib, dtype, shape):
    rng = np.random.default_rng(5)
    rows, cols = shape
    a = randn(rng, dtype, (rows, cols))
    if rows > 4:
        a[:, 2] = a[:, 0] - 2 * a[:, 1]  # rank deficient
    p, tau = fcopy(a), np.zeros(cols, dtype=dtype)
    nref = min(rows, cols)
    ncols = 3
    y0 = np.zeros((rows, ncols), dtype=dtype, order="F")
    y0[:nref] = randn(rng, dtype, (nref, ncols))
    y = fcopy(y0)
    hostlib.apply_host_qr(code(dtype), ptr(p), ctypes.c_int64(rows), cols, ptr(tau), ptr(y), ncols)
    r = np.triu(p[:nref, :])
    # Q from applying the reflectors to the identity
    eye = np.zeros((rows, nref), dtype=dtype, order="F")
    eye[:nref, :nref] = np.eye(nref)
    p2, tau2 = fcopy(a), np.zeros(cols, dtype=dtype)
    hostlib.apply_host_qr(code(dtype), ptr(p2), ctypes.c_int64(rows), cols, ptr(tau2), ptr(eye), nref)
    q = eye
    assert np.allclose(q.conj().T @ q, np.eye(nref), atol=1e-13)
    assert np.allclose(q @ r, a, atol=1e-13 * max(1.0, np.abs(a).max()) * rows)
    assert np.allclose(y, q @ y0[:nref], atol=1e-13 * rows)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("chi,rank", [(1, 1), (3, 3), (8, 8), (8, 5), (16, 16)])
def test_gauge_from_message(hostlib, dtype, chi, rank):
    rng = np.random.default_rng(chi * 10 + rank)
    f = randn(rng, dtype, (rank, chi))
    g = f.conj().T @ f                                   # Hermitian PSD of the given rank
    msg = fcopy(g + 1e-17 * randn(rng, dtype, (chi, chi)))  # not exactly Hermitian, like a BP message
    x = np.zeros((chi, chi), dtype=dtype, order="F")
    xinv = np.zeros_like(x)
    ev = np.zeros(chi)
    hostlib.apply_host_gauge(code(dtype), ptr(msg), chi, ptr(x), ptr(xinv), ev.ctypes.data_as(P))
    scale = np.abs(g).max()
    assert np.allclose(x.conj().T @ x, g, atol=1e-13 * scale)
    want_ev = np.linalg.eigvalsh(g)
    assert np.allclose(np.sort(ev), want_ev, atol=1e-13 * scale)
    proj = x @ xinv                                      # projector on the kept eigenvectors
    assert np.allclose(proj, np.diag(np.diag(proj)), atol=1e-10)
    assert int(round(np.trace(proj).real)) == rank
    # same gauge-invariant content as the oracle's factorisation
    xo, xinvo = A.gram_eigh_full_with_pinv(g)
    assert np.allclose(xinv @ x, xinvo @ xo, atol=1e-9)  # projector on the support, as an operator on the ket leg


def smem_for(sides, rb):
    """Emulated shared-memory budget (elements) that gives version 2 TSQR row blocks of `rb` (None: version 1)."""
    need = 0
    for z, d, slot, dims in sides:
        rows = int(np.prod([dims[i] for i in range(z) if i != slot], dtype=np.int64))
        cols = d * int(dims[slot])
        need = max(need, 2 * rows, cols * cols + 2 * cols * (cols + rb))
    return need


def two_site(hostlib, variant, dtype, a1, a2, op, max_rank, normalize, msg_out, sv):
    """variant: "v1" (csrc/bpx_apply.cuh) or ("v2", rb) (csrc/bpx_apply2.cuh with row blocks of rb)."""
    args = (code(dtype), a1[0], a1[1], a1[2], ptr(a1[3]), ptr(a1[4]), ptr(a1[5]), a2[0], a2[1], a2[2], ptr(a2[3]), ptr(a2[4]),
            ptr(a2[5]), ptr(op), int(max_rank), int(normalize), ptr(msg_out), sv.ctypes.data_as(P))
    if variant == "v1":
        return hostlib.apply_host_two_site(*args)
    if variant == "v3":
        return hostlib.apply_host_two_site_v3(*args)
    smem = smem_for([(a[0], a[1], a[2], a[3]) for a in (a1, a2)], variant[1])
    return hostlib.apply_host_two_site_v2(*args, ctypes.c_int64(smem))


VARIANTS = ["v1", ("v2", 1), ("v2", 3), ("v2", 8), ("v2", 4096), "v3"]


# ---------------------------------------------------------------------------------------------------------------
# whole gates against the oracle
# ---------------------------------------------------------------------------------------------------------------
def random_network(rng, dtype, adjacency, dims, d):
    """adjacency: {v: [neighbours in slot order]}, dims: {frozenset(v, w): chi}.  -> oracle state + PSD environment."""
    link = lambda v, w: ("l",) + tuple(sorted((v, w)))
    state, env = {}, {}
    for v, nb in adjacency.items():
        shape = (d[v],) + tuple(dims[frozenset((v, w))] for w in nb)
        names = (("s", v),) + tuple(link(v, w) for w in nb)
        state[v] = (randn(rng, dtype, shape), names)
        for w in nb:
            c = dims[frozenset((v, w))]
            f = randn(rng, dtype, (c, c)) + 1.5 * np.eye(c)
            m = f.conj().T @ f
            env[(w, v)] = (m / np.trace(m).real).astype(dtype)
    return state, env


def side_args(state, env, adjacency, v, other):
    x, names = state[v]
    nb = adjacency[v]
    dims = np.array(x.shape[1:], dtype=np.int32)
    slot = nb.index(other) if other is not None else -1
    msgs = np.concatenate([fcopy(env[(w, v)]).ravel(order="F") for w in nb]) if nb else np.zeros(0, x.dtype)
    return len(nb), x.shape[0], slot, dims, fcopy(x).ravel(order="F").copy(), np.ascontiguousarray(msgs)


def bond_product(state, v1, v2):
    """Gauge-invariant content of the pair: the two tensors contracted over their bond."""
    t = A.contract(state[v1], state[v2])
    names = sorted(t[1], key=repr)
    return A.permute(t, names)


GRID = {  # 2 x 3 grid, vertex -> neighbours in slot order
    0: [1, 3], 1: [0, 2, 4], 2: [1, 5], 3: [0, 4], 4: [3, 1, 5], 5: [4, 2],
}


def grid_dims(chi_bond, chi_other):
    dims = {}
    for v, nb in GRID.items():
        for w in nb:
            dims[frozenset((v, w))] = chi_other
    dims[frozenset((1, 4))] = chi_bond
    dims[frozenset((0, 1))] = max(1, chi_other - 1)  # ragged external legs
    return dims


@pytest.mark.parametrize("variant", VARIANTS, ids=str)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("chi_bond,chi_other,max_rank,normalize", [
    (3, 2, 0, False), (3, 2, 2, False), (4, 3, 3, True), (2, 3, 1, True), (4, 1, 0, False), (1, 2, 0, False),
    (2, 4, 2, True), (2, 8, 2, False),
])
def test_two_site_gate_matches_oracle(hostlib, variant, dtype, chi_bond, chi_other, max_rank, normalize):
    rng = np.random.default_rng(100 * chi_bond + 10 * chi_other + max_rank)
    d = {v: 2 for v in GRID}
    d[4] = 3  # different physical dims on the two sides
    state, env = random_network(rng, dtype, GRID, grid_dims(chi_bond, chi_other), d)
    v1, v2 = 1, 4
    o = randn(rng, dtype, (d[v1], d[v2], d[v1], d[v2]))
    names = (("s", v1), ("s", v2))
    # the device keeps the leg's dimension: max_rank = 0 means "truncate to chi_bond" (the fixed-chi simple update)
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=max_rank or chi_bond, normalize=normalize)

    a1 = side_args(state, env, GRID, v1, v2)
    a2 = side_args(state, env, GRID, v2, v1)
    (z1, d1, s1, dims1, site1, msgs1), (z2, d2, s2, dims2, site2, msgs2) = a1, a2
    op = fcopy(o).ravel(order="F").copy()
    msg_out = np.zeros(chi_bond * chi_bond, dtype=dtype)
    sv = np.zeros(chi_bond)
    rc = two_site(hostlib, variant, dtype, a1, a2, op, max_rank, normalize, msg_out, sv)
    assert rc == 0
    k = want_env[(v1, v2)].shape[0]
    s_want = np.diag(want_env[(v1, v2)]).real
    assert np.allclose(sv[:k], s_want, rtol=1e-10, atol=1e-13) and np.all(sv[k:] == 0)
    m = msg_out.reshape(chi_bond, chi_bond, order="F")
    assert np.allclose(m[:k, :k], want_env[(v1, v2)], rtol=1e-10, atol=1e-13)
    assert np.all(m[k:, :] == 0) and np.all(m[:, k:] == 0)

    got_state = dict(state)
    got_state[v1] = (site1.reshape(state[v1][0].shape, order="F"), state[v1][1])
    got_state[v2] = (site2.reshape(state[v2][0].shape, order="F"), state[v2][1])
    # the kept rank is zero-padded up to the leg's dimension
    b1 = got_state[v1][0].take(range(k, chi_bond), axis=1 + s1)
    assert b1.size == 0 or np.all(b1 == 0)
    got = bond_product(got_state, v1, v2)
    want = bond_product(want_state, v1, v2)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("variant", ["v1", ("v2", 1), ("v2", 64), "v3"], ids=str)
@pytest.mark.parametrize("dtype", DTYPES)
def test_two_site_gate_on_a_leaf_pair(hostlib, variant, dtype):
    """Both vertices have no external legs (a 2-site chain): rows = 1, no reflectors, no gauges."""
    rng = np.random.default_rng(2)
    adj = {0: [1], 1: [0]}
    state, env = random_network(rng, dtype, adj, {frozenset((0, 1)): 3}, {0: 2, 1: 2})
    o = randn(rng, dtype, (2, 2, 2, 2))
    names = (("s", 0), ("s", 1))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=3)
    a1, a2 = side_args(state, env, adj, 0, 1), side_args(state, env, adj, 1, 0)
    msg_out, sv = np.zeros(9, dtype=dtype), np.zeros(3)
    op = fcopy(o).ravel(order="F").copy()
    assert two_site(hostlib, variant, dtype, a1, a2, op, 0, 0, msg_out, sv) == 0
    got_state = {0: (a1[4].reshape(2, 3, order="F"), state[0][1]), 1: (a2[4].reshape(2, 3, order="F"), state[1][1])}
    k = want_env[(0, 1)].shape[0]                      # min(chi, m, n) = 1: rank of a 1 x d by d x 1 ... bond matrix
    assert np.allclose(sv[:k], np.diag(want_env[(0, 1)]).real, rtol=1e-12)
    assert np.abs(bond_product(got_state, 0, 1) - bond_product(want_state, 0, 1)).max() < 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("normalize", [False, True])
def test_one_site_gate_matches_oracle(hostlib, dtype, normalize):
    rng = np.random.default_rng(9)
    d = {v: 3 for v in GRID}
    state, env = random_network(rng, dtype, GRID, grid_dims(3, 2), d)
    v = 4
    o = randn(rng, dtype, (3, 3))
    names = (("s", v),)
    want_state, _ = A.apply_operator((o, names, names), state, env, normalize=normalize)
    z, dv, _, dims, site, msgs = side_args(state, env, GRID, v, None)
    op = fcopy(o).ravel(order="F").copy()
    hostlib.apply_host_one_site(code(dtype), z, dv, ptr(dims), ptr(site), ptr(msgs), ptr(op), int(normalize))
    got = site.reshape(state[v][0].shape, order="F")
    want = A.permute(want_state[v], state[v][1])
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("variant", ["v1", ("v2", 5), ("v2", 32), ("v2", 100000)], ids=str)
@pytest.mark.parametrize("dtype", DTYPES)
def test_two_site_gate_cfg5_shape(hostlib, variant, dtype):
    """The shape of BASELINE config 5's bulk (degree 4, chi = 16 would be 1 MiB per tensor: chi = 6 keeps the CPU test
    quick) with a rank-deficient incoming message."""
    rng = np.random.default_rng(77)
    adj = {0: [1, 2, 3, 4], 1: [0, 5, 6, 7]}
    for w in range(2, 8):
        adj[w] = [0 if w < 5 else 1]
    dims = {frozenset((v, w)): 6 for v, nb in adj.items() for w in nb}
    d = {v: 2 for v in adj}
    state, env = random_network(rng, dtype, adj, dims, d)
    f = randn(rng, dtype, (4, 6))
    env[(2, 0)] = (f.conj().T @ f).astype(dtype)       # rank 4 of 6: two gauge directions are dropped by the pinv
    o = randn(rng, dtype, (2, 2, 2, 2))
    names = (("s", 0), ("s", 1))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=6, normalize=True)
    a1, a2 = side_args(state, env, adj, 0, 1), side_args(state, env, adj, 1, 0)
    msg_out, sv = np.zeros(36, dtype=dtype), np.zeros(6)
    op = fcopy(o).ravel(order="F").copy()
    assert two_site(hostlib, variant, dtype, a1, a2, op, 6, 1, msg_out, sv) == 0
    assert np.allclose(sv, np.diag(want_env[(0, 1)]).real, rtol=1e-9)
    got_state = dict(state)
    got_state[0] = (a1[4].reshape(state[0][0].shape, order="F"), state[0][1])
    got_state[1] = (a2[4].reshape(state[1][0].shape, order="F"), state[1][1])
    got, want = bond_product(got_state, 0, 1), bond_product(want_state, 0, 1)
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()


@pytest.mark.parametrize("variant", ["v1", ("v2", 224), ("v2", 96), "v3"], ids=str)
def test_identity_gate_at_the_true_cfg5_shape(hostlib, variant):
    """BASELINE config 5's bulk shape for real: degree 4, chi = 16, d = 2 (1 MiB per tensor, 4096 x 32 matrix view).  The
    identity gate with the full rank kept must leave the pair invariant (checked through random probes on the external
    legs -- the full pair product would be 0.5 GB) and return the singular values of the gauged bond."""
    rng = np.random.default_rng(16)
    chi, d, z = 16, 2, 4
    dims = np.full(z, chi, dtype=np.int32)
    sites = [randn(rng, np.float64, (d,) + (chi,) * z) / 64.0 for _ in range(2)]
    msgs = []
    for _ in range(2):
        ms = []
        for _ in range(z):
            f = randn(rng, np.float64, (chi, chi)) + 4.0 * np.eye(chi)
            m = f.T @ f
            ms.append((m / np.trace(m)).ravel(order="F"))
        msgs.append(np.concatenate(ms))
    flat = [fcopy(s).ravel(order="F").copy() for s in sites]
    op = np.eye(d * d).reshape(d, d, d, d).ravel(order="F").copy()
    msg_out, sv = np.zeros(chi * chi), np.zeros(chi)
    slot1, slot2 = 1, 2
    a1 = (z, d, slot1, dims, flat[0], msgs[0])
    a2 = (z, d, slot2, dims, flat[1], msgs[1])
    rc = two_site(hostlib, variant, np.float64, a1, a2, op, 0, 0, msg_out, sv)
    assert rc == 0 and np.all(sv > 0) and np.all(np.diff(sv) <= 0)
    new = [f.reshape((d,) + (chi,) * z, order="F") for f in flat]

    def probe(t, slot, vecs):
        """Contract every external leg with a probe vector -> [s, bond]."""
        out = t
        for leg in reversed(range(z)):
            if leg != slot:
                out = np.tensordot(out, vecs[leg], axes=([1 + leg], [0]))
        return out

    v1 = [rng.standard_normal(chi) for _ in range(z)]
    v2 = [rng.standard_normal(chi) for _ in range(z)]
    before = probe(sites[0], slot1, v1) @ probe(sites[1], slot2, v2).T
    after = probe(new[0], slot1, v1) @ probe(new[1], slot2, v2).T
    assert np.abs(after - before).max() <= 1e-9 * np.abs(before).max()
    assert np.allclose(np.diag(msg_out.reshape(chi, chi, order="F")), sv)


# ---------------------------------------------------------------------------------------------------------------
# the other BASELINE shapes and numerically hard inputs
# ---------------------------------------------------------------------------------------------------------------
def star_pair(rng, dtype, z, chi, d, chi_bond=None, msg=None):
    """Two degree-z vertices joined by one bond, every other leg ending in a leaf; returns (adj, state, env)."""
    chi_bond = chi if chi_bond is None else chi_bond
    adj = {0: [1] + [2 + i for i in range(z - 1)], 1: [0] + [2 + (z - 1) + i for i in range(z - 1)]}
    for w in range(2, 2 * z):
        adj[w] = [0 if w < 2 + (z - 1) else 1]
    dims = {frozenset((v, w)): chi for v, nb in adj.items() for w in nb}
    dims[frozenset((0, 1))] = chi_bond
    state, env = random_network(rng, dtype, adj, dims, {v: d for v in adj})
    if msg is not None:
        for key in list(env):
            if key[1] in (0, 1) and key[0] not in (0, 1):
                env[key] = msg(rng, env[key].shape[0]).astype(dtype)
    return adj, state, env


def run_pair(hostlib, variant, dtype, adj, state, env, o, max_rank, normalize):
    a1, a2 = side_args(state, env, adj, 0, 1), side_args(state, env, adj, 1, 0)
    chi_b = int(a1[3][a1[2]])
    msg_out, sv = np.zeros(chi_b * chi_b, dtype=dtype), np.zeros(chi_b)
    op = fcopy(o).ravel(order="F").copy()
    assert two_site(hostlib, variant, dtype, a1, a2, op, max_rank, normalize, msg_out, sv) == 0
    got = dict(state)
    got[0] = (a1[4].reshape(state[0][0].shape, order="F"), state[0][1])
    got[1] = (a2[4].reshape(state[1][0].shape, order="F"), state[1][1])
    return got, sv


@pytest.mark.parametrize("variant", ["v1", ("v2", 32), ("v2", 100000), "v3"], ids=str)
@pytest.mark.parametrize("dtype,z,chi,d", [(np.float64, 6, 4, 2), (np.complex128, 3, 16, 2), (np.float64, 4, 8, 2),
                                          (np.complex128, 4, 8, 2)], ids=["cfg4", "cfg3", "cfg2", "cfg2c"])
def test_two_site_gate_baseline_shapes(hostlib, variant, dtype, z, chi, d):
    """The bulk shapes of BASELINE configs 2, 2c, 3 and 4 (config 5 and config 1 are covered above)."""
    rng = np.random.default_rng(z * 100 + chi)
    adj, state, env = star_pair(rng, dtype, z, chi, d)
    o = randn(rng, dtype, (d, d, d, d))
    names = (("s", 0), ("s", 1))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=chi, normalize=True)
    got, sv = run_pair(hostlib, variant, dtype, adj, state, env, o, chi, True)
    assert np.allclose(sv, np.diag(want_env[(0, 1)]).real, rtol=1e-9, atol=1e-13)
    x, y = bond_product(got, 0, 1), bond_product(want_state, 0, 1)
    assert np.abs(x - y).max() <= 1e-9 * np.abs(y).max()


def graded_message(decades):
    def make(rng, c):
        q, _ = np.linalg.qr(rng.standard_normal((c, c)) + 1j * rng.standard_normal((c, c)))
        ev = np.logspace(0, -decades, c)
        return (q * ev) @ q.conj().T
    return make


@pytest.mark.parametrize("variant", ["v1", ("v2", 16)], ids=str)
@pytest.mark.parametrize("decades", [6, 10, 13])
def test_two_site_gate_with_ill_conditioned_environments(hostlib, variant, decades):
    """Message spectra spanning 6, 10 and 13 decades (a strongly entangled bond next to a nearly product one): the gauges
    X = sqrt(D) V^H and their inverses amplify rounding by sqrt(cond).  Yardstick: the IDENTITY gate with the full rank
    kept must leave the pair invariant.  The device path (one-sided Jacobi: small eigenvalues of a PSD matrix to high
    relative accuracy) is at least as accurate as the oracle's LAPACK eigh there (measured: 1.6e-7 against 9.5e-4 at 13
    decades), and on a random gate the two agree to within their own errors; the singular values agree to 1e-8."""
    rng = np.random.default_rng(decades)
    dtype = np.complex128
    adj, state, env = star_pair(rng, dtype, 3, 4, 2, msg=graded_message(decades))
    names = (("s", 0), ("s", 1))
    ident = np.eye(4).reshape(2, 2, 2, 2).astype(dtype)
    y0 = bond_product(state, 0, 1)
    oracle_state, _ = A.apply_operator((ident, names, names), state, env, trunc=4)
    device_state, _ = run_pair(hostlib, variant, dtype, adj, state, env, ident, 4, False)
    err_oracle = np.abs(bond_product(oracle_state, 0, 1) - y0).max() / np.abs(y0).max()
    err_device = np.abs(bond_product(device_state, 0, 1) - y0).max() / np.abs(y0).max()
    assert err_device <= max(2 * err_oracle, 1e-10), (err_device, err_oracle)
    assert err_device <= {6: 1e-9, 10: 1e-5, 13: 1e-5}[decades]
    o = randn(rng, dtype, (2, 2, 2, 2))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=4, normalize=True)
    got, sv = run_pair(hostlib, variant, dtype, adj, state, env, o, 4, True)
    assert np.allclose(sv, np.diag(want_env[(0, 1)]).real, rtol=1e-8, atol=1e-12)
    x, y = bond_product(got, 0, 1), bond_product(want_state, 0, 1)
    assert np.abs(x - y).max() <= (20 * (err_oracle + err_device) + 1e-9) * np.abs(y).max()


@pytest.mark.parametrize("variant", ["v1", ("v2", 16), "v3"], ids=str)
def test_two_site_gate_with_degenerate_singular_values(hostlib, variant):
    """A gate that leaves an exactly degenerate spectrum on the bond (identity on a product state of Bell-like pairs):
    any basis of the degenerate subspace is a valid answer; the pair product and S are unique."""
    rng = np.random.default_rng(1)
    dtype = np.float64
    adj = {0: [1], 1: [0]}
    state = {0: (np.eye(2)[:, :] / np.sqrt(2), (("s", 0), ("l", 0, 1))), 1: (np.eye(2) / np.sqrt(2) * np.sqrt(2), (("s", 1), ("l", 0, 1)))}
    env = {(0, 1): np.eye(2) / 2, (1, 0): np.eye(2) / 2}
    o = np.eye(4).reshape(2, 2, 2, 2)
    names = (("s", 0), ("s", 1))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=2)
    got, sv = run_pair(hostlib, variant, dtype, adj, state, env, o, 2, False)
    assert np.allclose(sv, np.diag(want_env[(0, 1)]).real, rtol=1e-12) and abs(sv[0] - sv[1]) < 1e-14
    x, y = bond_product(got, 0, 1), bond_product(want_state, 0, 1)
    assert np.abs(x - y).max() <= 1e-13


@pytest.mark.parametrize("dtype", DTYPES)
def test_jacobi_svd_graded_spectrum(hostlib, dtype):
    """cfg5's bond matrix size (64 x 64) with singular values over 12 decades: every one to 1e-10 RELATIVE accuracy
    (one-sided Jacobi is relatively accurate; this is what keeps small Schmidt values meaningful under truncation)."""
    rng = np.random.default_rng(3)
    n = 64
    u, _ = np.linalg.qr(randn(rng, dtype, (n, n)))
    v, _ = np.linalg.qr(randn(rng, dtype, (n, n)))
    sv = np.logspace(0, -12, n)
    a = (u * sv) @ v.conj().T
    b, vv = fcopy(a), np.zeros((n, n), dtype=dtype, order="F")
    hostlib.apply_host_jacobi(code(dtype), ptr(b), n, n, ptr(vv))
    got = np.sort(np.linalg.norm(b, axis=0))[::-1]
    # the matrix itself carries eps * |A| of rounding from its construction: values above that are relatively accurate
    assert np.allclose(got[:40], sv[:40], rtol=1e-8)
    assert np.allclose(got, np.linalg.svd(a, compute_uv=False), atol=1e-15)
    assert np.allclose(vv.conj().T @ vv, np.eye(n), atol=1e-13)


# ---------------------------------------------------------------------------------------------------------------
# two-site expectation values (csrc/bpx_expect2.cuh)
# ---------------------------------------------------------------------------------------------------------------
def oracle_problem_from_state(oracle, adjacency, state):
    """Oracle `Problem` (bp_oracle.py) for a named state: vertices in dict order, legs already in slot order."""
    from itnn_b200.graphs import GraphArrays

    verts = list(adjacency)
    vi = {v: i for i, v in enumerate(verts)}
    src, dst, slot, row_ptr = [], [], [], [0]
    for v in verts:
        for k, w in enumerate(adjacency[v]):
            src.append(vi[v]); dst.append(vi[w]); slot.append(k)
        row_ptr.append(len(src))
    index = {(s, d): e for e, (s, d) in enumerate(zip(src, dst))}
    rev = [index[(d, s)] for s, d in zip(src, dst)]
    ga = GraphArrays(vertices=verts, vindex=vi, src=src, dst=dst, rev=rev, slot=slot, row_ptr=row_ptr, edge_index=index)
    return ga, oracle.make_problem(ga, [state[v][0] for v in verts], "norm")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("chi_bond,chi_other", [(3, 2), (2, 4), (1, 3), (4, 1)])
def test_edge_expect_matches_oracle(oracle, hostlib, dtype, chi_bond, chi_other):
    rng = np.random.default_rng(chi_bond * 10 + chi_other)
    d = {v: 2 for v in GRID}
    d[4] = 3
    state, env = random_network(rng, dtype, GRID, grid_dims(chi_bond, chi_other), d)
    ga, p = oracle_problem_from_state(oracle, GRID, state)
    msgs = [env[(ga.vertices[ga.src[e]], ga.vertices[ga.dst[e]])] for e in range(ga.ne)]
    v1, v2 = 1, 4
    o = randn(rng, dtype, (d[v1], d[v2], d[v1], d[v2]))
    want_num, want_den = oracle.two_site_expect(p, msgs, ga.edge_index[(ga.vindex[v1], ga.vindex[v2])], o)
    a1, a2 = side_args(state, env, GRID, v1, v2), side_args(state, env, GRID, v2, v1)
    num, den = np.zeros(1, dtype=dtype), np.zeros(1, dtype=dtype)
    op = fcopy(o).ravel(order="F").copy()
    rc = hostlib.apply_host_edge_expect(code(dtype), a1[0], a1[1], a1[2], ptr(a1[3]), ptr(a1[4]), ptr(a1[5]), a2[0], a2[1], a2[2],
                                        ptr(a2[3]), ptr(a2[4]), ptr(a2[5]), ptr(op), ptr(num), ptr(den))
    assert rc == 0
    assert np.isclose(num[0], want_num, rtol=1e-11, atol=1e-14 * abs(want_den))
    assert np.isclose(den[0], want_den, rtol=1e-11) and abs(np.imag(den[0])) <= 1e-12 * abs(den[0])


@pytest.mark.parametrize("dtype", DTYPES)
def test_edge_expect_is_exact_on_a_tree(oracle, hostlib, dtype):
    """Known answer: on a tree with converged BP messages the two-site expectation value is the exact one."""
    rng = np.random.default_rng(8)
    adj = {0: [1], 1: [0, 2, 3], 2: [1], 3: [1, 4], 4: [3]}
    dims = {frozenset((v, w)): 3 for v, nb in adj.items() for w in nb}
    state, _ = random_network(rng, dtype, adj, dims, {v: 2 for v in adj})
    ga, p = oracle_problem_from_state(oracle, adj, state)
    msgs = [np.ones((3, 3), dtype=dtype) for _ in range(ga.ne)]
    for _ in range(6):  # diameter 3: the synchronous schedule has converged
        msgs = oracle.sweep_jacobi(p, msgs)
    env = {(ga.vertices[ga.src[e]], ga.vertices[ga.dst[e]]): msgs[e] for e in range(ga.ne)}
    psi = A.permute(A.prod(state), [("s", v) for v in adj]).reshape(-1)
    for v1, v2 in ((1, 3), (3, 4), (0, 1)):
        o = randn(rng, dtype, (2, 2, 2, 2))
        a1, a2 = side_args(state, env, adj, v1, v2), side_args(state, env, adj, v2, v1)
        num, den = np.zeros(1, dtype=dtype), np.zeros(1, dtype=dtype)
        op = fcopy(o).ravel(order="F").copy()
        hostlib.apply_host_edge_expect(code(dtype), a1[0], a1[1], a1[2], ptr(a1[3]), ptr(a1[4]), ptr(a1[5]), a2[0], a2[1], a2[2],
                                       ptr(a2[3]), ptr(a2[4]), ptr(a2[5]), ptr(op), ptr(num), ptr(den))
        full = psi.reshape((2,) * 5)
        moved = np.moveaxis(full, (v1, v2), (0, 1))
        opsi = np.tensordot(o, moved, axes=([2, 3], [0, 1]))
        exact = np.vdot(moved.ravel(), opsi.ravel()) / np.vdot(psi, psi)
        assert np.isclose(num[0] / den[0], exact, rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------
# version 3 (the Gram path, csrc/bpx_apply3.cuh): what it declines, and how accurate it is where it does not
# ---------------------------------------------------------------------------------------------------------------
def run_pair_v3(hostlib, dtype, adj, state, env, o, max_rank, normalize):
    """-> (status, new state or None, singular values); a declined gate (status 1) must leave every input untouched."""
    a1, a2 = side_args(state, env, adj, 0, 1), side_args(state, env, adj, 1, 0)
    before = (a1[4].copy(), a2[4].copy())
    chi_b = int(a1[3][a1[2]])
    msg_out, sv = np.zeros(chi_b * chi_b, dtype=dtype), np.zeros(chi_b)
    op = fcopy(o).ravel(order="F").copy()
    rc = two_site(hostlib, "v3", dtype, a1, a2, op, max_rank, normalize, msg_out, sv)
    assert rc in (0, 1)
    if rc == 1:
        assert np.array_equal(a1[4], before[0]) and np.array_equal(a2[4], before[1]) and not msg_out.any()
        return 1, None, sv
    got = dict(state)
    got[0] = (a1[4].reshape(state[0][0].shape, order="F"), state[0][1])
    got[1] = (a2[4].reshape(state[1][0].shape, order="F"), state[1][1])
    return 0, got, sv


@pytest.mark.parametrize("dtype", DTYPES)
def test_v3_declines_a_rank_deficient_message(hostlib, dtype):
    """The reference projects on the support of a rank-deficient message (pinv, apply_operators.jl:250-253); the Gram
    shortcut assumes X^-1 X = 1, so version 3 must hand such a gate back untouched (the caller re-runs it on version 2)."""
    rng = np.random.default_rng(77)
    adj, state, env = star_pair(rng, dtype, 4, 4, 2)
    f = randn(rng, dtype, (3, 4))
    env[(2, 0)] = (f.conj().T @ f).astype(dtype)  # rank 3 of 4
    o = randn(rng, dtype, (2, 2, 2, 2))
    rc, got, _ = run_pair_v3(hostlib, dtype, adj, state, env, o, 4, True)
    assert rc == 1
    # an indefinite message is declined too (the reference keeps its positive part only)
    adj, state, env = star_pair(rng, dtype, 3, 3, 2)
    env[(2, 0)] = np.diag([1.0, 0.5, -0.2]).astype(dtype)
    rc, got, _ = run_pair_v3(hostlib, dtype, adj, state, env, o, 3, False)
    assert rc == 1


@pytest.mark.parametrize("decades", [2, 4, 6, 10, 13])
def test_v3_on_ill_conditioned_environments(hostlib, decades):
    """Graded message spectra: version 3 either declines the gate (conditioning of the Gram matrix, COND_MIN) or is as
    accurate as the identity-gate yardstick demands (same bounds as versions 1 / 2 above)."""
    rng = np.random.default_rng(decades)
    dtype = np.complex128
    adj, state, env = star_pair(rng, dtype, 3, 4, 2, msg=graded_message(decades))
    names = (("s", 0), ("s", 1))
    ident = np.eye(4).reshape(2, 2, 2, 2).astype(dtype)
    y0 = bond_product(state, 0, 1)
    rc, got, _ = run_pair_v3(hostlib, dtype, adj, state, env, ident, 4, False)
    if decades >= 10:
        assert rc == 1  # three messages of 10 decades each: far beyond what a Gram matrix resolves
        return
    if rc == 1:
        return
    err = np.abs(bond_product(got, 0, 1) - y0).max() / np.abs(y0).max()
    assert err <= {2: 1e-12, 4: 1e-10, 6: 1e-8}[decades], err
    o = randn(rng, dtype, (2, 2, 2, 2))
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=4, normalize=True)
    rc, got, sv = run_pair_v3(hostlib, dtype, adj, state, env, o, 4, True)
    assert rc == 0
    assert np.allclose(sv, np.diag(want_env[(0, 1)]).real, rtol=1e-8, atol=1e-12)
    x, y = bond_product(got, 0, 1), bond_product(want_state, 0, 1)
    assert np.abs(x - y).max() <= 1e-7 * np.abs(y).max()


def test_v3_fuzz(hostlib):
    """Random shapes like test_two_site_gate_fuzz: version 3 takes every well-conditioned case and agrees with the oracle."""
    rng = np.random.default_rng(4048)
    worst, applied = 0.0, 0
    for it in range(80):
        dtype = DTYPES[it % 2]
        z, chi, chib, d = int(rng.integers(1, 5)), int(rng.integers(1, 5)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
        adj, state, env = star_pair(rng, dtype, z, chi, d, chi_bond=chib)
        o = randn(rng, dtype, (d, d, d, d))
        k, norm = int(rng.integers(1, chib + 1)), bool(rng.integers(0, 2))
        names = (("s", 0), ("s", 1))
        want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=k, normalize=norm)
        s_want = np.diag(want_env[(0, 1)]).real
        y = bond_product(want_state, 0, 1)
        rc, got, sv = run_pair_v3(hostlib, dtype, adj, state, env, o, k, norm)
        if rc == 0:
            applied += 1
            worst = max(worst, np.abs(sv[:len(s_want)] - s_want).max() / s_want.max(),
                        np.abs(bond_product(got, 0, 1) - y).max() / np.abs(y).max())
    assert applied >= 76, applied  # random_network's messages are well conditioned (F^H F with F = randn + 1.5 I)
    assert worst < 1e-9, worst


def test_two_site_gate_fuzz(hostlib):
    """120 random shapes (degree 1-4, link dims 1-4, bond 1-5, d 1-3, random truncation / normalisation, both dtypes),
    versions 1 and 2 with random TSQR block heights, against the oracle (an offline run of 800 cases: worst 1.2e-11)."""
    rng = np.random.default_rng(2024)
    worst = 0.0
    for it in range(60):
        dtype = DTYPES[it % 2]
        z, chi, chib, d = int(rng.integers(1, 5)), int(rng.integers(1, 5)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
        adj, state, env = star_pair(rng, dtype, z, chi, d, chi_bond=chib)
        o = randn(rng, dtype, (d, d, d, d))
        k, norm = int(rng.integers(1, chib + 1)), bool(rng.integers(0, 2))
        names = (("s", 0), ("s", 1))
        want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=k, normalize=norm)
        s_want = np.diag(want_env[(0, 1)]).real
        y = bond_product(want_state, 0, 1)
        for variant in ("v1", ("v2", int(rng.integers(1, 20)))):
            got, sv = run_pair(hostlib, variant, dtype, adj, state, env, o, k, norm)
            worst = max(worst, np.abs(sv[:len(s_want)] - s_want).max() / s_want.max(),
                        np.abs(bond_product(got, 0, 1) - y).max() / np.abs(y).max())
    assert worst < 1e-9, worst


def test_edge_expect_fuzz(oracle, hostlib):
    """60 random shapes (degree 1-4, link dims 1-4, bond 1-5, d 1-3, both dtypes) against oracle.two_site_expect."""
    rng = np.random.default_rng(77)
    for it in range(60):
        dtype = DTYPES[it % 2]
        z, chi, chib, d = int(rng.integers(1, 5)), int(rng.integers(1, 5)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
        adj, state, env = star_pair(rng, dtype, z, chi, d, chi_bond=chib)
        ga, p = oracle_problem_from_state(oracle, adj, state)
        msgs = [env[(ga.vertices[ga.src[e]], ga.vertices[ga.dst[e]])] for e in range(ga.ne)]
        o = randn(rng, dtype, (d, d, d, d))
        want_num, want_den = oracle.two_site_expect(p, msgs, ga.edge_index[(0, 1)], o)
        a1, a2 = side_args(state, env, adj, 0, 1), side_args(state, env, adj, 1, 0)
        num, den = np.zeros(1, dtype=dtype), np.zeros(1, dtype=dtype)
        op = fcopy(o).ravel(order="F").copy()
        hostlib.apply_host_edge_expect(code(dtype), a1[0], a1[1], a1[2], ptr(a1[3]), ptr(a1[4]), ptr(a1[5]), a2[0], a2[1], a2[2],
                                       ptr(a2[3]), ptr(a2[4]), ptr(a2[5]), ptr(op), ptr(num), ptr(den))
        assert np.isclose(den[0], want_den, rtol=1e-11)
        assert np.isclose(num[0], want_num, rtol=1e-10, atol=1e-13 * abs(want_den))


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "c128"])
@pytest.mark.parametrize("z,chi,chib,d", [(4, 6, 5, 2), (4, 7, 4, 2), (3, 13, 6, 2), (3, 15, 3, 3)],
                         ids=["rows216", "rows343", "rows169", "rows225_d3"])
def test_v3_row_counts_that_end_in_a_partial_tile(hostlib, dtype, z, chi, chib, d):
    """Matrix views of 169 .. 343 rows: one or two full 128-row tiles of the Gram and final passes followed by a partial one
    (the scratch copies are column-group major: a tile is one piece per group, `copy_tile_groups`), odd row counts, and an
    odd column count (d = 3)."""
    rng = np.random.default_rng(z * 100 + chi)
    adj, state, env = star_pair(rng, dtype, z, chi, d, chi_bond=chib)
    o = randn(rng, dtype, (d, d, d, d))
    names = (("s", 0), ("s", 1))
    k = max(1, chib - 1)
    want_state, want_env = A.apply_operator((o, names, names), state, env, trunc=k, normalize=True)
    s_want = np.diag(want_env[(0, 1)]).real
    y = bond_product(want_state, 0, 1)
    rc, got, sv = run_pair_v3(hostlib, dtype, adj, state, env, o, k, True)
    assert rc == 0
    assert np.abs(sv[:len(s_want)] - s_want).max() <= 1e-9 * s_want.max()
    assert np.abs(bond_product(got, 0, 1) - y).max() <= 1e-9 * np.abs(y).max()
