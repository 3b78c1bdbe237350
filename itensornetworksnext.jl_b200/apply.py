"""Host-side mirror of the reference's gate application, backed by libbpx (bpx_apply_*_gates, CUDA).

Same names, argument meaning and error behaviour as /root/reference/src/apply/apply_operators.jl:
  apply_operators(operators, state, env; alg, operator_alg, environment_alg, kwargs...)  ............ :28-60
  apply_operator(operator, state, env; alg, kwargs...) -> (state, env)  ............................ :166-176
  BPApplyGate(trunc, normalize) -- the default strategy ............................................ :180-211
  apply_gate_bp! (one- and two-site) ............................................................... :213-283
  NoApplyOperatorEnvironmentPreparation ............................................................ :131-146
`state` is an ITensorNetwork (the ket), `env` a MessageCache of operator-shaped messages M_e[bra, ket] keyed by directed
edge (what `message_environment` / `beliefpropagation` on the NormNetwork return).  Inputs are not modified
(`initialize_output` copies, :204-208).

ALL arithmetic (message eigen-decompositions, gauging, QR, the gate, the truncated SVD, inverse gauges) runs in the CUDA
library; this module lowers names to the canonical layout, calls the C ABI and wraps the result.  Nothing here falls
back to the CPU and oracle/ is never imported.

Beyond the reference: `apply_operators` sends every run of consecutive vertex-disjoint gates (a circuit layer) to the
device in ONE call -- the order of disjoint gates does not matter, so the result is the reference's.
The device keeps a link's dimension fixed during a call; when the reference would change the bond dimension (no `trunc`
on a bond that can grow, or `trunc` below the current dimension) the bond is re-declared (zero padded) before the call
and sliced afterwards, so the returned tensors have exactly the reference's dimensions.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from .beliefpropagation import AbstractAlgorithm, ArgumentError, MessageCache
from .device import BPXContext, cast_to
from .graphs import NamedEdge
from .tensornetwork import Index, ITensor, ITensorNetwork, NormNetwork, canonical_arrays


class Operator:
    """`operator(array, out_names, in_names)` of ITensorBase: axes (out..., in...) (test/test_apply_operator.jl:19-27)."""

    def __init__(self, data, out_names: Sequence[Hashable], in_names: Sequence[Hashable]):
        self.data = np.asarray(data)
        self.out_names = tuple(out_names)
        self.in_names = tuple(in_names)
        if self.data.ndim != len(self.out_names) + len(self.in_names) or len(self.out_names) != len(self.in_names):
            raise ValueError("operator axes must be (out..., in...) with as many outputs as inputs")

    def domainnames(self):
        return self.in_names


class ApplyOperatorAlgorithm(AbstractAlgorithm):
    """apply_operators.jl:150."""


@dataclass
class BPApplyGate(ApplyOperatorAlgorithm):
    """apply_operators.jl:180-183.  `trunc`: None or the rank kept by the SVD (`truncrank`)."""

    trunc: Optional[int] = None
    normalize: bool = False
    device: int = 0


class NoApplyOperatorEnvironmentPreparation(AbstractAlgorithm):
    """apply_operators.jl:131-139: the environments are not touched between operators."""


def _default_operator_algorithm(**kwargs) -> BPApplyGate:
    try:
        return BPApplyGate(**kwargs)
    except TypeError as e:  # Julia: MethodError for an unknown keyword
        raise ArgumentError(str(e)) from None


def _select_operator_algorithm(alg, **kwargs) -> ApplyOperatorAlgorithm:
    if alg is None:
        return _default_operator_algorithm(**kwargs)
    if isinstance(alg, dict):
        if kwargs:
            raise ArgumentError("Additional keyword arguments are not allowed when `alg` is a `NamedTuple`.")
        return _default_operator_algorithm(**alg)
    if isinstance(alg, AbstractAlgorithm):
        if kwargs:
            raise ArgumentError("Additional keyword arguments are not allowed when `alg` is an `AbstractAlgorithm` instance.")
        return alg
    raise TypeError(f"cannot select an algorithm from {alg!r}")


# ---------------------------------------------------------------------------------------------------------------
# lowering
# ---------------------------------------------------------------------------------------------------------------
def _touched_vertices(op: Operator, state: ITensorNetwork) -> List:
    ins = set(op.domainnames())
    vs = [v for v in state.vertices() if ins & set(state.sitenames(v))]
    if not vs:
        raise ArgumentError("operator shares no indices with the tensor network")
    if len(vs) > 2:
        raise ArgumentError(f"{len(vs)}-site gate decomposition not implemented")
    return vs


def _lowered_operator(op: Operator, state: ITensorNetwork, vs: Sequence, dtype) -> np.ndarray:
    """-> op[o_1, (o_2,) i_1, (i_2)] with the site legs of every vertex fused like `canonical_arrays` fuses them."""
    sites = [state.sitenames(v) for v in vs]
    flat = [n for s in sites for n in s]
    if set(flat) != set(op.in_names) or len(flat) != len(op.in_names):
        raise ArgumentError("an operator must act on all site indices of the vertices it touches")
    out_of = dict(zip(op.in_names, range(len(op.in_names))))          # in name -> position among the outputs
    axes = [out_of[n] for n in flat] + [len(flat) + op.in_names.index(n) for n in flat]
    arr = np.transpose(op.data, axes)
    dims = [int(np.prod([state[v].data.shape[state[v].dimnames().index(n)] for n in s], dtype=np.int64)) for v, s in zip(vs, sites)]
    return cast_to(np.ascontiguousarray(arr).reshape(dims + dims), dtype, "operator")


def _promoted(state: ITensorNetwork, ops: Sequence[Operator], env) -> ITensorNetwork:
    """ITensorBase.apply promotes: a complex gate (exp(-i dt H)) or complex messages on a Float64 state give a ComplexF64
    state.  The device session takes its dtype from the state, so promote the state here instead of truncating the gate."""
    if np.dtype(state.dtype).kind == "c":
        return state
    need = any(np.iscomplexobj(op.data) for op in ops)
    if not need and env is not None:
        need = any(np.iscomplexobj(m.data if isinstance(m, ITensor) else m) for m in env.values())
    if not need:
        return state
    return ITensorNetwork({v: ITensor(t.data.astype(np.complex128), t.inds) for v, t in state.tensors.items()})


class _ApplySession:
    """state + env resident on one GPU in the canonical layout (optionally with one bond zero-padded)."""

    def __init__(self, state: ITensorNetwork, env: Optional[MessageCache], device: int, pad: Optional[Tuple[int, int]] = None):
        self.state = state
        self.nn = NormNetwork(state, {n: ("bra", n) for n, vs in state.dimname_vertices.items() if len(vs) == 2})
        cp = canonical_arrays(self.nn)
        self.cp, ga = cp, cp.ga
        msgs = []
        for e in range(ga.ne):
            ne_ = ga.named_edge(e)
            if env is None:  # all-ones messages (test/test_apply_operator.jl:72)
                msgs.append(np.ones((cp.link_dim[e], cp.link_dim[e]), dtype=cp.dtype))
                continue
            if ne_ not in env:
                raise KeyError(f"no message on edge {ne_!r}")
            msgs.append(self._message_array(env[ne_], cp.ket_names[e], cp.dtype))
        tensors, link_dim = list(cp.tensors), list(cp.link_dim)
        if pad is not None:  # grow one bond to `new_dim` with zeros (an equivalent state)
            e, new_dim = pad
            r = ga.rev[e]
            for ee in (e, r):
                v, slot, old = ga.src[ee], ga.slot[ee], link_dim[ee]
                widths = [(0, 0)] * tensors[v].ndim
                widths[1 + slot] = (0, new_dim - old)
                tensors[v] = np.pad(tensors[v], widths)
                msgs[ee] = np.pad(msgs[ee], ((0, new_dim - old), (0, new_dim - old)))
                link_dim[ee] = new_dim
        self.link_dim = link_dim
        self.shapes = [t.shape for t in tensors]
        self.ctx = BPXContext(device)
        self.ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        self.ctx.set_dims(cp.dtype, "norm", cp.phys_dim, link_dim)
        self.ctx.set_site_tensors(tensors)
        self.ctx.set_messages(msgs)

    @staticmethod
    def _message_array(m, ket_name, dtype) -> np.ndarray:
        """env[e] -> [bra, ket]: the ket axis is the one that carries the state's link name."""
        if isinstance(m, ITensor):
            names = m.dimnames()
            if len(names) != 2 or ket_name not in names:
                raise ArgumentError(f"message {names} is not an operator on link {ket_name!r}")
            bra = names[0] if names[1] == ket_name else names[1]
            return cast_to(m.array(bra, ket_name), dtype, "message")
        return cast_to(m, dtype, "message")

    def site_itensor(self, v, keep: Optional[Dict[int, int]] = None) -> ITensor:
        """Download vertex `v` and name its axes like the input state (site legs un-fused; `keep`: slot -> kept dim)."""
        ga, net = self.cp.ga, self.state
        vi = ga.vindex[v]
        arr = self.ctx.get_site_tensor(vi).reshape(self.shapes[vi], order="F")
        old = net[v]
        sites = net.sitenames(v)
        links = [net.linkname(NamedEdge(v, w)) for w in net.graph.neighbors(v)]
        for slot, k in (keep or {}).items():
            arr = np.take(arr, range(k), axis=1 + slot)
        by = {i.name: i for i in old.inds}
        site_inds = [by[n] for n in sites]
        link_inds = [Index(arr.shape[1 + j], n) for j, n in enumerate(links)]
        arr = arr.reshape(tuple(i.dim for i in site_inds) + arr.shape[1:])
        return ITensor(np.ascontiguousarray(arr), site_inds + link_inds)

    def close(self):
        self.ctx.close()


def _reference_rank(cp, e: int, trunc: Optional[int]) -> int:
    """The bond dimension the reference's `svd_trunc` leaves on edge e (apply_operators.jl:258-261)."""
    ga = cp.ga
    bound = []
    for ee in (e, ga.rev[e]):
        v, slot = ga.src[ee], ga.slot[ee]
        shape = cp.tensors[v].shape
        rows = int(np.prod([shape[1 + j] for j in range(len(shape) - 1) if j != slot], dtype=np.int64))
        cols = shape[0] * shape[1 + slot]
        bound.append(min(rows, cols) * shape[0])
    k = min(bound)
    return k if trunc is None else min(int(trunc), k)


def _diag_message(old, s: np.ndarray, ket_name, dtype) -> ITensor:
    k = len(s)
    if isinstance(old, ITensor):
        names = old.dimnames()
        bra = names[0] if names[1] == ket_name else names[1]
    else:
        bra = ("bra", ket_name)
    return ITensor(np.diag(s).astype(dtype), (Index(k, bra), Index(k, ket_name)))


# ---------------------------------------------------------------------------------------------------------------
# apply_operator / apply_operators
# ---------------------------------------------------------------------------------------------------------------
def _apply_batch(alg: BPApplyGate, ops: Sequence[Operator], touched: Sequence[Sequence], state: ITensorNetwork,
                 env: MessageCache):
    """A run of vertex-disjoint gates of one kind (all one-site or all two-site) in one device call."""
    state = _promoted(state, ops, env)
    tensors = dict(state.tensors)
    new_env = env.copy()
    two = len(touched[0]) == 2
    probe = canonical_arrays(NormNetwork(state, {n: ("bra", n) for n, vs in state.dimname_vertices.items() if len(vs) == 2}))
    pad = None
    ranks = []
    if two:
        for vs in touched:
            if vs[1] not in state.graph.neighbors(vs[0]):
                raise ArgumentError(f"two-site gate on vertices {vs[0]!r}, {vs[1]!r} that share no link")
            e = probe.ga.edge_id(NamedEdge(vs[0], vs[1]))
            k = _reference_rank(probe, e, alg.trunc)
            ranks.append((e, k))
            if k > probe.link_dim[e]:
                if len(touched) > 1:
                    raise AssertionError("bond-growing gates are applied one at a time")  # guaranteed by the caller
                pad = (e, k)
    s = _ApplySession(state, env, alg.device, pad)
    try:
        dtype = s.cp.dtype
        lowered = [_lowered_operator(op, state, vs, dtype) for op, vs in zip(ops, touched)]
        if two:
            # one max_rank per call: gates whose reference rank differs from the common one are handled by slicing
            max_rank = max(k for _, k in ranks)
            svs = s.ctx.apply_two_site_gates([e for e, _ in ranks], lowered, max_rank=max_rank, normalize=alg.normalize)
            ga = s.cp.ga
            for (e, k), vs, sv in zip(ranks, touched, svs):
                r = ga.rev[e]
                tensors[vs[0]] = s.site_itensor(vs[0], {ga.slot[e]: k})
                tensors[vs[1]] = s.site_itensor(vs[1], {ga.slot[r]: k})
                ket = s.cp.ket_names[e]
                e12, e21 = NamedEdge(vs[0], vs[1]), NamedEdge(vs[1], vs[0])
                new_env[e12] = _diag_message(env[e12], sv[:k], ket, dtype)
                new_env[e21] = _diag_message(env[e21], sv[:k], ket, dtype)
        else:
            s.ctx.apply_one_site_gates([s.cp.ga.vindex[vs[0]] for vs in touched], lowered, normalize=alg.normalize)
            for vs in touched:
                tensors[vs[0]] = s.site_itensor(vs[0])
    finally:
        s.close()
    return ITensorNetwork(tensors), new_env


def apply_operator(operator: Operator, state: ITensorNetwork, env, alg=None, **kwargs):
    """-> (state, env) with `operator` applied (apply_operators.jl:166-176); only the messages on the gate edge change."""
    algorithm = _select_operator_algorithm(alg, **kwargs)
    if not isinstance(algorithm, BPApplyGate):
        raise TypeError(f"unsupported apply_operator algorithm {type(algorithm).__name__}")
    env = env if isinstance(env, MessageCache) else MessageCache(env)
    vs = _touched_vertices(operator, state)
    return _apply_batch(algorithm, [operator], [vs], state, env)


def apply_operators(operators: Sequence[Operator], state: ITensorNetwork, env, alg=None, operator_alg=None,
                    environment_alg=None, **kwargs):
    """Operators applied in turn (apply_operators.jl:28-60, 106-121); no environment preparation between them."""
    if alg is not None:
        raise ArgumentError("only the default apply_operators iteration is mirrored; pass `operator_alg=` / kwargs")
    if environment_alg is not None and not isinstance(environment_alg, NoApplyOperatorEnvironmentPreparation):
        raise TypeError(f"unsupported environment preparation {type(environment_alg).__name__}")
    algorithm = _select_operator_algorithm(operator_alg, **kwargs)
    if not isinstance(algorithm, BPApplyGate):
        raise TypeError(f"unsupported apply_operator algorithm {type(algorithm).__name__}")
    env = env if isinstance(env, MessageCache) else MessageCache(env)
    operators = list(operators)
    if not operators:
        return ITensorNetwork(dict(state.tensors)), env.copy()  # :55
    i = 0
    while i < len(operators):
        vs0 = _touched_vertices(operators[i], state)
        batch_ops, batch_vs, used = [operators[i]], [vs0], set(vs0)
        # extend the run while the next gates are of the same kind, vertex-disjoint and keep their bond dimension
        if not (len(vs0) == 2 and _changes_bond(state, vs0, algorithm.trunc)):
            j = i + 1
            while j < len(operators):
                vs = _touched_vertices(operators[j], state)
                if len(vs) != len(vs0) or used & set(vs) or (len(vs) == 2 and _changes_bond(state, vs, algorithm.trunc)):
                    break
                batch_ops.append(operators[j])
                batch_vs.append(vs)
                used |= set(vs)
                j += 1
        state, env = _apply_batch(algorithm, batch_ops, batch_vs, state, env)
        i += len(batch_ops)
    return state, env


def _changes_bond(state: ITensorNetwork, vs: Sequence, trunc: Optional[int]) -> bool:
    """Would the reference leave a different dimension on the gate bond than it has now?"""
    if vs[1] not in state.graph.neighbors(vs[0]):
        return True  # reported by the single-gate path
    bound = []
    for v, w in ((vs[0], vs[1]), (vs[1], vs[0])):
        t = state[v]
        link = state.linkname(NamedEdge(v, w))
        sites = set(state.sitenames(v))
        d = chi = rows = 1
        for i in t.inds:
            if i.name in sites:
                d *= i.dim
            elif i.name == link:
                chi = i.dim
            else:
                rows *= i.dim
        bound.append(min(rows, d * chi) * d)
    k = min(bound) if trunc is None else min(int(trunc), min(bound))
    return k != chi


def expect_two_site(operators: Sequence[Operator], state: ITensorNetwork, env, device: int = 0) -> List:
    """<O> of two-site operators on neighbouring vertices of `state` in the BP environment `env` (bpx_edge_expect).
    Build-defined like `expect` (the reference has neither); read-only, the operators may overlap."""
    env = env if isinstance(env, MessageCache) else MessageCache(env)
    operators = list(operators)
    if not operators:
        return []
    state = _promoted(state, operators, env)
    s = _ApplySession(state, env, device)
    try:
        edges, lowered = [], []
        for op in operators:
            vs = _touched_vertices(op, state)
            if len(vs) != 2 or vs[1] not in state.graph.neighbors(vs[0]):
                raise ArgumentError("expect_two_site takes operators on two neighbouring vertices")
            edges.append(s.cp.ga.edge_id(NamedEdge(vs[0], vs[1])))
            lowered.append(_lowered_operator(op, state, vs, s.cp.dtype))
        num, den = s.ctx.edge_expect(edges, lowered)
        return list(num / den)
    finally:
        s.close()
