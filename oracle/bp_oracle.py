"""CPU ORACLE (numpy) for the belief-propagation message-update path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product path (`itensornetworksnext.jl_b200`) never does; it fails loudly when the
CUDA library is missing.

This is a restatement, in plain numpy (`einsum`/`tensordot`, float64 / complex128), of the reference's
algorithm for the path named by BASELINE.json `north_star`.  Reference citations (relative to
/root/reference):

  * per-edge update ........ src/beliefpropagation/beliefpropagation.jl:242-257
        gather incoming messages (all in-edges of src(e) except reverse(e),
        src/beliefpropagation/messagecache.jl:124-131), contract them with the factor,
        divide by `sum(new_message)` unless that sum is exactly zero, replace cache[e].
  * NormNetwork factor ..... src/normnetwork.jl:49-54, 66-82   (ket ⊗ conj(bra), site leg shared)
  * message axes ........... src/beliefpropagation/messagecache.jl:205-225 (bra link, ket link)
  * sweep .................. src/beliefpropagation/beliefpropagation.jl:200-210 (in place, edge list)
  * residual ............... src/beliefpropagation/beliefpropagation.jl:261-267
  * stop logic ............. src/AlgorithmsInterfaceExtensions/AlgorithmsInterfaceExtensions.jl:84-119,
                             src/beliefpropagation/beliefpropagation.jl:16-55
  * beliefs ................ src/beliefpropagation/messagecache.jl:139-201

The arithmetic of the reference executes in un-vendored Julia dependencies (ITensorBase 0.10,
TensorAlgebra 0.16); exact contraction is order independent
(test/test_contract_network.jl:13-45), so the restatement needs only the tensor-network semantics.

PINNING.  The reference holds no golden vectors for this path and Julia is absent, so the oracle is
pinned by the reference's own known-answer tests (tests/test_oracle_known_answers.py):
tree exactness in one sequential sweep (test/test_beliefpropagation.jl:157-202), spin-ice
z = 1.5^(n^2) (test/test_beliefpropagation.jl:204-225), iterate_diff(c, copy(c)) = 0 (:134-150), the
incoming-message exclusion rule (:104-114), <psi|psi> of a NormNetwork (test/test_normnetwork.jl:148-166),
and by agreement with an independent C restatement (oracle/bp_oracle.c) and with the literal
double-layer contraction.  NOT pinned by any reference test ("parity unpinned" for these): per-sweep
message values on a loopy NormNetwork, local expectation values, and the synchronous (Jacobi) schedule,
which the reference does not have (its sweep is sequential, beliefpropagation.jl:200-210, 255) and which
is expressed here in reference terms as "for each e: c = copy(old); message_update!(c, e); new[e] = c[e]".

Canonical data (same as include/bpx.h):
  graph ....... directed edges e = 0..ne-1 grouped by source; src[e], dst[e], rev[e], slot[e];
                row_ptr[v]..row_ptr[v+1] are the out-edges of v, in link-leg order.
  norm mode ... site tensor A_v as ndarray of shape (d, chi_0, ..., chi_{z-1}) (legs in out-edge order);
                message on edge e as ndarray (chi, chi) indexed [bra, ket].
  single mode . factor T_v of shape (chi_0, ..., chi_{z-1}); message a vector (chi,).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


@dataclass
class Problem:
    src: np.ndarray
    dst: np.ndarray
    rev: np.ndarray
    slot: np.ndarray
    row_ptr: np.ndarray
    tensors: List[np.ndarray]
    mode: str = "norm"  # "norm" (double layer) | "single"

    @property
    def nv(self) -> int:
        return len(self.row_ptr) - 1

    @property
    def ne(self) -> int:
        return len(self.src)

    def degree(self, v: int) -> int:
        return int(self.row_ptr[v + 1] - self.row_ptr[v])

    def out_edges(self, v: int) -> range:
        return range(int(self.row_ptr[v]), int(self.row_ptr[v + 1]))


def make_problem(ga, tensors: Sequence[np.ndarray], mode: str = "norm") -> Problem:
    """`ga` is anything with src/dst/rev/slot/row_ptr sequences (e.g. graphs.GraphArrays)."""
    return Problem(
        np.asarray(ga.src, dtype=np.int64),
        np.asarray(ga.dst, dtype=np.int64),
        np.asarray(ga.rev, dtype=np.int64),
        np.asarray(ga.slot, dtype=np.int64),
        np.asarray(ga.row_ptr, dtype=np.int64),
        [np.asarray(t) for t in tensors],
        mode,
    )


# ---------------------------------------------------------------------------------------------------
# one message update  (beliefpropagation.jl:242-257)
# ---------------------------------------------------------------------------------------------------
def incoming_edges(p: Problem, e: int) -> List[Optional[int]]:
    """Per link leg of src(e): the edge whose message flows INTO src(e) on that leg; None for e's own leg.

    messagecache.jl:128-131 -- in_incident_edges(src(e)) minus reverse(e).
    """
    u = int(p.src[e])
    out = []
    for f in p.out_edges(u):
        out.append(None if f == e else int(p.rev[f]))
    return out


def normalize_message(m: np.ndarray) -> np.ndarray:
    """beliefpropagation.jl:248-253: divide by the SUM of all entries unless it is exactly zero."""
    s = m.sum()
    if s != 0:
        m = m / s
    return m


def contract_norm(A: np.ndarray, slot: int, in_msgs: Sequence[Optional[np.ndarray]]) -> np.ndarray:
    """Mtilde[b', b] = sum A[s, .., b, ..] conj(A[s, .., b', ..]) prod_i M_i[a'_i, a_i]   (absorption order).

    `in_msgs[i]` is the message arriving on link leg i ([bra, ket]); entry `slot` is ignored.
    """
    z = A.ndim - 1
    T = A
    for i in range(z):
        if i == slot:
            continue
        M = in_msgs[i]
        # T'[.., a', ..] = sum_a M[a', a] T[.., a, ..]
        T = np.moveaxis(np.tensordot(M, T, axes=([1], [i + 1])), 0, i + 1)
    # close with the bra layer: sum over s and all primed legs except `slot`
    axes = [ax for ax in range(z + 1) if ax != slot + 1]
    # result[b (from T), b' (from conj A)] -> transpose to [bra, ket] = [b', b]
    res = np.tensordot(T, np.conj(A), axes=(axes, axes))
    return res.T.copy()


def contract_norm_literal(A: np.ndarray, slot: int, in_msgs: Sequence[Optional[np.ndarray]]) -> np.ndarray:
    """The reference's literal evaluation: materialise the ket ⊗ conj(bra) factor
    (chi^(2z) elements; normnetwork.jl:49-54 makes it an atomic leaf of the contraction) and then
    contract the messages into it.  Only feasible for small chi, z."""
    z = A.ndim - 1
    ket = _LETTERS[1:1 + z]
    bra = _LETTERS[1 + z:1 + 2 * z]
    D = np.einsum(f"a{ket},a{bra}->{ket}{bra}", A, np.conj(A))  # double-layer factor, site leg summed
    operands, subs = [D], [ket + bra]
    for i in range(z):
        if i == slot:
            continue
        operands.append(in_msgs[i])
        subs.append(bra[i] + ket[i])
    out = bra[slot] + ket[slot]
    return np.einsum(",".join(subs) + "->" + out, *operands)


def contract_single(T: np.ndarray, slot: int, in_msgs: Sequence[Optional[np.ndarray]]) -> np.ndarray:
    """Single-layer factor: m~[b] = sum_alpha T[.., b, ..] prod_i m_i[alpha_i]."""
    z = T.ndim
    X = T
    # contract from the last leg down so axis numbers of earlier legs stay valid
    for i in reversed(range(z)):
        if i == slot:
            continue
        X = np.tensordot(X, in_msgs[i], axes=([i], [0]))
    return np.asarray(X).reshape(T.shape[slot])


def message_update(p: Problem, msgs: Sequence[np.ndarray], e: int, normalize: bool = True,
                   literal: bool = False) -> np.ndarray:
    u = int(p.src[e])
    ins = [None if f is None else msgs[f] for f in incoming_edges(p, e)]
    if p.mode == "norm":
        fn = contract_norm_literal if literal else contract_norm
        new = fn(p.tensors[u], int(p.slot[e]), ins)
    else:
        new = contract_single(p.tensors[u], int(p.slot[e]), ins)
    if normalize:
        new = normalize_message(new)
    return new


# ---------------------------------------------------------------------------------------------------
# sweeps
# ---------------------------------------------------------------------------------------------------
def sweep_sequential(p: Problem, msgs: List[np.ndarray], edge_seq: Sequence[int], normalize: bool = True,
                     literal: bool = False) -> List[np.ndarray]:
    """The reference schedule: in-place updates along an explicit edge list (beliefpropagation.jl:200-210, 255)."""
    msgs = list(msgs)
    for e in edge_seq:
        msgs[e] = message_update(p, msgs, int(e), normalize, literal)
    return msgs


def sweep_jacobi(p: Problem, msgs: Sequence[np.ndarray], normalize: bool = True, literal: bool = False,
                 edges: Optional[Sequence[int]] = None) -> List[np.ndarray]:
    """Synchronous schedule: every directed edge updated from the previous sweep's messages."""
    new = list(msgs)
    for e in (range(p.ne) if edges is None else edges):
        new[e] = message_update(p, msgs, int(e), normalize, literal)
    return new


# ---------------------------------------------------------------------------------------------------
# residual and stopping  (beliefpropagation.jl:261-267, AIE.jl:84-119)
# ---------------------------------------------------------------------------------------------------
def edge_residual(m1: np.ndarray, m2: np.ndarray) -> float:
    n1 = np.linalg.norm(m1.ravel())
    n2 = np.linalg.norm(m2.ravel())
    d = np.vdot(m1.ravel() / n1, m2.ravel() / n2)
    return float(1.0 - abs(d) ** 2)


def iterate_diff(msgs1: Sequence[np.ndarray], msgs2: Sequence[np.ndarray]) -> float:
    return max(edge_residual(a, b) for a, b in zip(msgs1, msgs2))


def beliefpropagation(p: Problem, msgs: Sequence[np.ndarray], maxiter: Optional[int] = None,
                      tol: Optional[float] = None, schedule: str = "jacobi",
                      edge_seq: Optional[Sequence[int]] = None, normalize: bool = True,
                      history: Optional[list] = None):
    """Outer loop with the reference's stop semantics: the criterion is evaluated before every sweep,
    the residual is ignored at iteration 0, stop when residual < tol or iteration >= maxiter.
    Returns (messages, iterations_done, last_residual)."""
    if maxiter is None and tol is None:
        raise ValueError("At least one of `maxiter` or `tol` must be specified.")
    msgs = [np.array(m) for m in msgs]
    it, delta = 0, float("inf")
    prev = msgs
    while True:
        if it > 0:
            delta = iterate_diff(msgs, prev)
            if history is not None:
                history.append(delta)
            if tol is not None and delta < tol:
                break
        if maxiter is not None and it >= maxiter:
            break
        prev = msgs
        if schedule == "jacobi":
            msgs = sweep_jacobi(p, msgs, normalize)
        elif schedule == "sequential":
            msgs = sweep_sequential(p, msgs, edge_seq, normalize)
        else:
            raise ValueError(schedule)
        it += 1
    return msgs, it, delta


# ---------------------------------------------------------------------------------------------------
# beliefs  (messagecache.jl:139-201)
# ---------------------------------------------------------------------------------------------------
def vertex_scalar(p: Problem, msgs: Sequence[np.ndarray], v: int, op: Optional[np.ndarray] = None):
    """Factor at v contracted with ALL incoming messages (messagecache.jl:139-143).

    `op` (d x d, indexed [s_out, s_in]) optionally acts on the ket site leg: the numerator of a local
    expectation value (build-defined extension; the reference has no `expect`)."""
    ins = [msgs[int(p.rev[f])] for f in p.out_edges(v)]
    A = p.tensors[v]
    if p.mode == "single":
        X = A
        for i in reversed(range(A.ndim)):
            X = np.tensordot(X, ins[i], axes=([i], [0]))
        return np.asarray(X).reshape(())[()]
    z = A.ndim - 1
    T = A
    for i in range(z):
        T = np.moveaxis(np.tensordot(ins[i], T, axes=([1], [i + 1])), 0, i + 1)
    if op is not None:
        T = np.tensordot(op, T, axes=([1], [0]))
    return np.vdot(A.ravel(), T.ravel())  # sum conj(A) * T


def vertex_scalars(p: Problem, msgs: Sequence[np.ndarray]):
    return [vertex_scalar(p, msgs, v) for v in range(p.nv)]


def edge_scalar(p: Problem, msgs: Sequence[np.ndarray], e: int):
    """contract(cache[e], cache[reverse(e)]) (messagecache.jl:153-157): the two messages share both
    names, so this is the plain (unconjugated) sum of elementwise products."""
    return (msgs[e] * msgs[int(p.rev[e])]).sum()


def edge_scalars(p: Problem, msgs: Sequence[np.ndarray]):
    """One scalar per undirected edge, in order of first appearance (messagecache.jl:161-178)."""
    return [edge_scalar(p, msgs, e) for e in range(p.ne) if e < int(p.rev[e])]


def bethe_free_energy(p: Problem, msgs: Sequence[np.ndarray]):
    """sum log(vertex scalars) - sum log(edge scalars), complex-promoted when a real part is negative,
    -inf when an edge scalar is zero (messagecache.jl:185-201)."""
    num = np.asarray(vertex_scalars(p, msgs))
    den = np.asarray(edge_scalars(p, msgs))
    if np.any(num.real < 0):
        num = num.astype(np.complex128)
    if np.any(den.real < 0):
        den = den.astype(np.complex128)
    if np.any(den == 0):
        return -np.inf
    return np.sum(np.log(num)) - np.sum(np.log(den))


def local_expect(p: Problem, msgs: Sequence[np.ndarray], v: int, op: np.ndarray):
    """<O_v> = vertex contraction with O on the ket site leg / vertex_scalar (build-defined, see above)."""
    return vertex_scalar(p, msgs, v, op) / vertex_scalar(p, msgs, v)


def edge_density(p: Problem, msgs: Sequence[np.ndarray], e: int) -> np.ndarray:
    """BP two-site reduced density matrix (unnormalised) of the vertices of directed edge e, rho[s1, s2, s1', s2'] (ket
    indices first): both norm-network factors (normnetwork.jl:49-54) with every incoming message except the two on the
    shared link -- the two-vertex analogue of `vertex_scalar` (messagecache.jl:139-143).  Build-defined, like
    `local_expect` (the reference has no `expect`)."""
    assert p.mode == "norm"

    def half(vtx, slot):
        A = p.tensors[vtx]
        z = A.ndim - 1
        ins = [msgs[int(p.rev[f])] for f in p.out_edges(vtx)]
        T = A
        for i in range(z):
            if i != slot:
                T = np.moveaxis(np.tensordot(ins[i], T, axes=([1], [i + 1])), 0, i + 1)  # M[bra, ket] on the ket leg
        ext = [i + 1 for i in range(z) if i != slot]
        return np.tensordot(T, A.conj(), axes=(ext, ext))  # N[s, b, s', b']

    n1 = half(int(p.src[e]), int(p.slot[e]))
    n2 = half(int(p.dst[e]), int(p.slot[int(p.rev[e])]))
    return np.einsum("abcd,ebfd->aecf", n1, n2)


def two_site_expect(p: Problem, msgs: Sequence[np.ndarray], e: int, op: np.ndarray):
    """(numerator, denominator) of <O_e>, op[o1, o2, i1, i2] with 1 = src(e), 2 = dst(e)."""
    rho = edge_density(p, msgs, e)
    return np.einsum("cdab,abcd->", op, rho), np.einsum("abab->", rho)


# ---------------------------------------------------------------------------------------------------
# brute force (for known-answer tests only)
# ---------------------------------------------------------------------------------------------------
def contract_all(p: Problem):
    """Exact contraction of the whole network by one einsum (small graphs only).

    single mode: the scalar prod(tn).  norm mode: <psi|psi> = ||prod(ket)||^2 (test/test_normnetwork.jl:148-166)."""
    link_id = {}
    operands = []
    site_labels = []
    for v in range(p.nv):
        labels = []
        if p.mode == "norm":
            site_labels.append(len(site_labels))
            labels.append(site_labels[-1])
        for f in p.out_edges(v):
            key = min(f, int(p.rev[f]))
            if key not in link_id:
                link_id[key] = p.nv + len(link_id)
            labels.append(link_id[key])
        operands += [p.tensors[v], labels]
    if p.mode == "norm":
        psi = np.einsum(*operands, site_labels, optimize="greedy")
        return np.vdot(psi.ravel(), psi.ravel())
    return np.einsum(*operands, [], optimize="greedy")[()]


def contract_all_sequential(p: Problem):
    """Exact contraction of a SINGLE-layer network by absorbing the vertices one at a time into a running tensor
    (vertex order; open legs = links to vertices not absorbed yet).  Same value as `contract_all`, but the cost is
    bounded by the cut width of the vertex order (a 4x4 periodic grid of chi = 2 keeps <= 2^10 entries), where one
    big einsum call can take minutes."""
    assert p.mode == "single"
    cur = np.ones((), dtype=np.result_type(*[t.dtype for t in p.tensors])) if p.nv else np.ones(())
    labels: List[int] = []
    for v in range(p.nv):
        mine = [min(f, int(p.rev[f])) for f in p.out_edges(v)]
        shared = [l for l in mine if l in labels]
        cur = np.tensordot(cur, p.tensors[v], axes=([labels.index(l) for l in shared], [mine.index(l) for l in shared]))
        labels = [l for l in labels if l not in shared] + [l for l in mine if l not in shared]
    assert not labels
    return cur[()]

