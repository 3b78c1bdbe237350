"""CPU ORACLE (numpy) for the main CONSUMER of BP messages: the BP simple-update gate application.
TEST INFRASTRUCTURE ONLY (same rules as bp_oracle.py: only tests/ may import it).

Restates /root/reference/src/apply/apply_operators.jl:
  * apply_gate_bp!  ................ :213-224  (which vertices an operator touches)
  * one-site gate .................. :226-244  (apply, optional normalisation by the gauged norm)
  * two-site gate .................. :246-283  (gauges from the incoming boundary messages -> QR -> gate -> truncated
                                                SVD -> sqrt(S) split -> inverse gauges; new messages diag(S) on the gate edge)
  * apply_operators (a sequence) ... :28-60, :106-121

Why it is here: the reference's apply_operator tests (test/test_apply_operator.jl:62-133) are the
only reference tests that pin BP messages on a *NormNetwork*: a truncated two-site gate on an open chain reproduces the
globally optimal truncated SVD if and only if the messages are the true environments, with the [bra, ket] orientation
right.  tests/test_apply_oracle.py runs those known-answer tests on the oracle's BP messages (CPU) and on the CUDA
path's messages (GPU).

Un-vendored dependencies restated from their contracts (the code is not under /root/reference):
  * TensorAlgebra.MatrixAlgebra.gram_eigh_full(_with_pinv)(G): for a Hermitian positive semi-definite G = V D V^H, the
    factor X = sqrt(D) V^H with X^H X = G, and its pseudo-inverse V D^{-1/2} (eigenvalues below a relative cutoff are
    dropped from the inverse).  apply_operators.jl:247-252 conjugates the factors because messages are stored as
    operators [bra, ket] while the gauges act on ket legs; in matrix terms the ket leg is multiplied by X.
  * MatrixAlgebraKit.qr_compact / svd_trunc(truncrank(k)): thin QR; SVD keeping the k largest singular values.

Data model: a state is {vertex: (ndarray, names)} with one name per axis; two tensors sharing a name are linked.
An environment is {(w, v): M} for every directed edge w -> v, M[bra, ket] over the link between v and w.
An operator is (ndarray, out_names, in_names) with axes (out..., in...); applying it contracts `in` with the state's
axes of those names and gives the result the same names again (ITensorBase.apply's gate semantics,
test/test_apply_operator.jl:22-27).
"""
from __future__ import annotations

import itertools
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

Named = Tuple[np.ndarray, Tuple[Hashable, ...]]
_fresh = itertools.count()


def fresh(prefix: str = "o") -> str:
    return f"{prefix}#{next(_fresh)}"


# ---------------------------------------------------------------------------------------------------
# named-array helpers
# ---------------------------------------------------------------------------------------------------
def contract(a: Named, b: Named) -> Named:
    """Product of two named arrays: shared names are summed over (ITensor `*`)."""
    (x, xn), (y, yn) = a, b
    shared = [n for n in xn if n in yn]
    out = np.tensordot(x, y, axes=([xn.index(n) for n in shared], [yn.index(n) for n in shared]))
    return out, tuple(n for n in xn if n not in shared) + tuple(n for n in yn if n not in shared)


def permute(a: Named, names: Sequence[Hashable]) -> np.ndarray:
    x, xn = a
    assert set(names) == set(xn) and len(names) == len(xn), (names, xn)
    return np.transpose(x, [xn.index(n) for n in names])


def rename(a: Named, mapping: Dict[Hashable, Hashable]) -> Named:
    x, xn = a
    return x, tuple(mapping.get(n, n) for n in xn)


def prod(state: Dict[Hashable, Named]) -> Named:
    """`prod(network)`: the full contraction, open (site) names left over."""
    it = iter(state.values())
    acc = next(it)
    for t in it:
        acc = contract(acc, t)
    return acc


def apply_op(op, a: Named) -> Named:
    """ITensorBase.apply(op, a): contract the operator's input names with `a`, name the outputs like the inputs."""
    o, out_names, in_names = op
    k = len(out_names)
    tmp_names = tuple(fresh("op") for _ in range(k))
    res = contract((o, tmp_names + tuple(in_names)), a)
    return rename(res, dict(zip(tmp_names, in_names)))


def neighbors(state: Dict[Hashable, Named], v) -> List:
    names = set(state[v][1])
    return [w for w in state if w != v and names & set(state[w][1])]


def linkname(state: Dict[Hashable, Named], v, w):
    shared = [n for n in state[v][1] if n in state[w][1]]
    assert len(shared) == 1, f"vertices {v!r}, {w!r} share {len(shared)} names"
    return shared[0]


def sitenames(state: Dict[Hashable, Named], v) -> List:
    others = set()
    for w in state:
        if w != v:
            others |= set(state[w][1])
    return [n for n in state[v][1] if n not in others]


# ---------------------------------------------------------------------------------------------------
# gauges from messages
# ---------------------------------------------------------------------------------------------------
def gram_eigh_full_with_pinv(m: np.ndarray, rtol: Optional[float] = None):
    """G = m [bra, ket], Hermitian PSD.  -> (X, Xinv): X^H X = G, X Xinv = projector on the support of G."""
    g = 0.5 * (m + m.conj().T)
    d, v = np.linalg.eigh(g)
    if rtol is None:
        rtol = np.finfo(d.dtype).eps * len(d)
    dmax = max(d.max(), 0.0)
    keep = d > rtol * dmax
    sq = np.sqrt(np.where(keep, d, 0.0))
    x = sq[:, None] * v.conj().T
    inv_sq = np.zeros_like(d)
    inv_sq[keep] = 1.0 / sq[keep]
    return x, v * inv_sq[None, :]


def _gauged(state, env, v, exclude=()):
    """state[v] with the gauge of every incoming boundary message applied to the corresponding ket leg.
    Returns (tensor, [(gauged name, original link name, Xinv)])."""
    t = state[v]
    undo = []
    for w in neighbors(state, v):
        if w in exclude:
            continue
        l = linkname(state, v, w)
        x, xinv = gram_eigh_full_with_pinv(env[(w, v)])
        g = fresh("g")
        t = contract(t, (x, (g, l)))          # ket leg l -> gauged leg g
        undo.append((g, l, xinv))
    return t, undo


def _ungauge(t: Named, undo) -> Named:
    for g, l, xinv in undo:
        t = contract(t, (xinv, (l, g)))       # X Xinv = 1 on the support
    return t


# ---------------------------------------------------------------------------------------------------
# apply_operator (BPApplyGate, apply_operators.jl:190-283)
# ---------------------------------------------------------------------------------------------------
def apply_operator(op, state: Dict[Hashable, Named], env: Dict[Tuple, np.ndarray], trunc: Optional[int] = None,
                   normalize: bool = False):
    """-> (new state, new env); inputs are not modified (`initialize_output` copies, :204-208)."""
    o, out_names, in_names = op
    vs = [v for v in state if set(in_names) & set(sitenames(state, v))]
    if not vs:
        raise ValueError("operator shares no indices with the tensor network")
    state, env = dict(state), dict(env)
    if len(vs) == 1:
        (v,) = vs
        psi = apply_op(op, state[v])
        if normalize:
            probe = dict(state)
            probe[v] = psi
            gauged, _ = _gauged(probe, env, v)
            psi = (psi[0] / np.linalg.norm(gauged[0].ravel()), psi[1])
        state[v] = psi
        return state, env
    if len(vs) != 2:
        raise ValueError(f"{len(vs)}-site gate decomposition not implemented")
    v1, v2 = vs
    bond = linkname(state, v1, v2)
    g1, undo1 = _gauged(state, env, v1, exclude=(v2,))
    g2, undo2 = _gauged(state, env, v2, exclude=(v1,))

    def qr_compact(t: Named, other: Named):
        rows = [n for n in t[1] if n not in other[1] and n not in in_names]   # :261-262
        cols = [n for n in t[1] if n not in rows]
        a = permute(t, rows + cols)
        rshape, cshape = a.shape[:len(rows)], a.shape[len(rows):]
        q, r = np.linalg.qr(a.reshape(int(np.prod(rshape, dtype=np.int64)), -1))
        qn = fresh("qr")
        return (q.reshape(rshape + (q.shape[1],)), tuple(rows) + (qn,)), (r.reshape((r.shape[0],) + cshape), (qn,) + tuple(cols))

    q1, r1 = qr_compact(g1, g2)
    q2, r2 = qr_compact(g2, g1)
    theta = apply_op(op, contract(r1, r2))
    rows = [n for n in r1[1] if n not in r2[1]]                                # :265
    cols = [n for n in theta[1] if n not in rows]
    a = permute(theta, rows + cols)
    rshape, cshape = a.shape[:len(rows)], a.shape[len(rows):]
    u, s, vh = np.linalg.svd(a.reshape(int(np.prod(rshape, dtype=np.int64)), -1), full_matrices=False)
    k = len(s) if trunc is None else min(int(trunc), len(s))
    u, s, vh = u[:, :k], s[:k], vh[:k, :]
    if normalize:
        s = s / np.linalg.norm(s)
    sq = np.sqrt(s)
    new1 = ((u * sq[None, :]).reshape(rshape + (k,)), tuple(rows) + (bond,))     # the new link keeps the old name
    new2 = ((sq[:, None] * vh).reshape((k,) + cshape), (bond,) + tuple(cols))
    state[v1] = _ungauge(contract(q1, new1), undo1)
    state[v2] = _ungauge(contract(q2, new2), undo2)
    env[(v1, v2)] = np.diag(s).astype(o.dtype if np.iscomplexobj(o) else s.dtype)  # :277-282
    env[(v2, v1)] = env[(v1, v2)].copy()
    return state, env


def apply_operators(ops, state, env, **kwargs):
    """Operators applied in turn; no environment preparation in between (NoApplyOperatorEnvironmentPreparation, :131-146)."""
    state, env = dict(state), dict(env)
    for op in ops:
        state, env = apply_operator(op, state, env, **kwargs)
    return state, env
