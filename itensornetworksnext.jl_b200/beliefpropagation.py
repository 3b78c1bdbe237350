"""Host-side mirror of the reference's belief-propagation interface, backed by libbpx (CUDA, sm_100a).

Same names, argument meaning and error behaviour as /root/reference:
  beliefpropagation(factors, messages; edges, stopping_criterion, message_update_algorithm)
        ........................ src/beliefpropagation/beliefpropagation.jl:69-92
  stopping-criterion shorthand .. :16-55      StopWhenConverged .. AlgorithmsInterfaceExtensions.jl:63-119
  MessageUpdateAlgorithm / SimpleMessageUpdate / message_update!  :214-257
  select_algorithm / default_algorithm .. src/select_algorithm.jl:9-51
  MessageCache, messagecache, message_environment, incoming_messages, vertex_scalar(s), edge_scalar(s),
  region_scalar, bethe_free_energy ........ src/beliefpropagation/messagecache.jl:15-229
  iterate_diff ........................... beliefpropagation.jl:261-267

Julia is not available in this image, so this module plays the role of julia/BPX.jl: it lowers the
reference's containers to the canonical layout and calls the C ABI.  ALL arithmetic of the path (message
updates, normalisation, residuals, vertex/edge scalars) runs in the CUDA library; nothing here falls
back to the CPU, and the oracle under oracle/ is never imported.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Hashable, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from . import algorithmsinterface as AI
from .algorithmsinterface import StopAfterIteration, StopWhenAny, StopWhenConverged, StoppingCriterion
from .device import BPXContext, cast_to
from .graphs import NamedEdge, forest_cover_edge_sequence, to_edge
from .tensornetwork import CanonicalProblem, Index, ITensor, ITensorNetwork, NormNetwork, canonical_arrays


class ArgumentError(ValueError):
    """Julia `ArgumentError` analogue."""


# ---------------------------------------------------------------------------------------------------
# stopping criteria (AlgorithmsInterface.StopAfterIteration, AIE.StopWhenConverged, `|`): algorithmsinterface.py
# ---------------------------------------------------------------------------------------------------
def select_beliefpropagation_stopping_criterion(c=None, **kwargs) -> StoppingCriterion:
    """beliefpropagation.jl:16-55."""
    if isinstance(c, StoppingCriterion):
        return c
    if c is None and not kwargs:
        raise ArgumentError(
            "`stopping_criterion` must be specified, e.g.\n"
            "  `stopping_criterion = dict(maxiter = 10)`,\n"
            "  `stopping_criterion = dict(maxiter = 10, tol = 1.0e-10)`, or\n"
            "  `stopping_criterion = StopAfterIteration(10) | StopWhenConverged(1.0e-10)`."
        )
    if c is not None:
        if not isinstance(c, dict):
            raise ArgumentError(f"unsupported `stopping_criterion` {c!r}")
        kwargs = dict(c)
    maxiter = kwargs.pop("maxiter", None)
    tol = kwargs.pop("tol", None)
    if kwargs:
        raise ArgumentError(f"Unrecognized `stopping_criterion` kwargs: {tuple(kwargs)}. Supported: `maxiter`, `tol`.")
    if maxiter is None and tol is None:
        raise ArgumentError("At least one of `maxiter` or `tol` must be specified.")
    crit = None
    if maxiter is not None:
        crit = StopAfterIteration(int(maxiter))
    if tol is not None:
        conv = StopWhenConverged(tol)
        crit = conv if crit is None else crit | conv
    return crit


class _NotFlat(Exception):
    """The criterion is not an OR of StopAfterIteration / StopWhenConverged: run the generic AI loop."""


def _flatten_criterion(c: StoppingCriterion):
    """-> (maxiter or None, tol or None) of an OR-combination."""
    maxiter = tol = None
    items = c.criteria if isinstance(c, StopWhenAny) else [c]
    for it in items:
        if isinstance(it, StopWhenAny):
            m, t = _flatten_criterion(it)
        elif isinstance(it, StopAfterIteration):
            m, t = it.maxiter, None
        elif isinstance(it, StopWhenConverged):
            m, t = None, it.tol
        else:
            raise _NotFlat()
        if m is not None:
            maxiter = m if maxiter is None else min(maxiter, m)
        if t is not None:
            tol = t if tol is None else max(tol, t)
    return maxiter, tol


# ---------------------------------------------------------------------------------------------------
# algorithm selection (src/select_algorithm.jl)
# ---------------------------------------------------------------------------------------------------
class AbstractAlgorithm:
    pass


class MessageUpdateAlgorithm(AbstractAlgorithm):
    """Strategy interface of beliefpropagation.jl:214-220."""


@dataclass
class SimpleMessageUpdate(MessageUpdateAlgorithm):
    """The reference's default strategy (beliefpropagation.jl:237-257): per-edge exact contraction with
    sum-normalisation, applied IN PLACE along the edge sequence (sequential schedule).  Here every
    update of the sequence runs in the CUDA library (bpx_sweep_sequence); runs of independent updates
    are batched into one launch."""

    normalize: bool = True
    device: int = 0
    schedule: str = "sequential"


@dataclass
class B200MessageUpdate(MessageUpdateAlgorithm):
    """The B200-native strategy: whole synchronous sweeps (every directed edge updated from the previous
    sweep's messages) in one C-ABI call, residual fused into the kernel epilogue (bpx_sweep)."""

    normalize: bool = True
    device: int = 0
    schedule: str = "synchronous"  # or "sequential"
    kernel: int = _lib.BPX_KERNEL_AUTO


def default_algorithm(f, args=None, **kwargs) -> AbstractAlgorithm:
    if f is message_update:
        return SimpleMessageUpdate(**kwargs)  # beliefpropagation.jl:223-225
    raise TypeError(f"no default algorithm for {getattr(f, '__name__', f)!r}")  # MethodError analogue


def select_algorithm(f, alg, args=None, **kwargs) -> AbstractAlgorithm:
    """src/select_algorithm.jl:16-51: None -> default; dict (NamedTuple) -> kwargs of the default;
    an AbstractAlgorithm instance is passed through untouched (the plugin point)."""
    if alg is None:
        return default_algorithm(f, args, **kwargs)
    if isinstance(alg, dict):
        if kwargs:
            raise ArgumentError("Additional keyword arguments are not allowed when `alg` is a `NamedTuple`.")
        return default_algorithm(f, args, **alg)
    if isinstance(alg, AbstractAlgorithm):
        if kwargs:
            raise ArgumentError(
                "Additional keyword arguments are not allowed when `alg` is an `AbstractAlgorithm` instance."
            )
        return alg
    raise TypeError(f"cannot select an algorithm from {alg!r}")


# ---------------------------------------------------------------------------------------------------
# MessageCache (messagecache.jl:15-120)
# ---------------------------------------------------------------------------------------------------
class MessageCache:
    """Directed edge -> message.  Values are `ITensor`s for BP; any value type is storable (the
    reference's container tests use strings)."""

    def __init__(self, messages=None):
        self._m: Dict[NamedEdge, object] = {}
        self._session: Optional["_Session"] = None
        self._session_version = -1  # version of the session's device message set that equals this cache's values
        if messages is not None:
            items = messages.items() if hasattr(messages, "items") else messages
            for e, m in items:
                self._m[to_edge(e)] = m

    # dictionary interface
    def __getitem__(self, e):
        return self._m[to_edge(e)]

    def __setitem__(self, e, m):
        self._m[to_edge(e)] = m
        self._session = None  # device copy is stale

    def __contains__(self, e):
        return to_edge(e) in self._m

    def __len__(self):
        return len(self._m)

    def keys(self):
        return self._m.keys()

    def values(self):
        return self._m.values()

    def items(self):
        return self._m.items()

    def edges(self):
        return list(self._m.keys())

    def has_edge(self, e):
        return to_edge(e) in self._m

    def vertices(self):
        seen = {}
        for e in self._m:
            seen.setdefault(e.src)
            seen.setdefault(e.dst)
        return list(seen)

    def copy(self) -> "MessageCache":
        c = MessageCache()
        c._m = {e: (m.copy() if hasattr(m, "copy") else m) for e, m in self._m.items()}
        return c

    def copyto(self, src, edges: Optional[Iterable] = None) -> "MessageCache":
        items = src.items() if hasattr(src, "items") else src
        if edges is not None:
            edges = {to_edge(e) for e in edges}
        for e, m in items:
            if edges is None or to_edge(e) in edges:
                self[e] = m
        return self

    def map(self, f: Callable) -> "MessageCache":
        return MessageCache({e: f(m) for e, m in self._m.items()})

    def in_incident_edges(self, v) -> List[NamedEdge]:
        return [e for e in self._m if e.dst == v]

    def subgraph(self, vertices) -> "MessageCache":
        vs = set(vertices)
        return MessageCache({e: m for e, m in self._m.items() if e.src in vs and e.dst in vs})

    def iterate_diff(self, other) -> float:
        """`AIE.iterate_diff(::MessageCache, ::MessageCache)` (beliefpropagation.jl:261-267), on the device."""
        return iterate_diff(self, other)


def messagecache(f_or_pairs, edges=None) -> MessageCache:
    if edges is None:
        return MessageCache(dict(f_or_pairs))
    return MessageCache({to_edge(e): f_or_pairs(to_edge(e)) for e in edges})


def incoming_messages(cache: MessageCache, edge) -> List:
    """All messages into src(edge) except the one along reverse(edge) (messagecache.jl:124-131)."""
    e = to_edge(edge)
    return [cache[f] for f in cache.in_incident_edges(e.src) if f != e.reverse()]


def similar_message_environment(nn: NormNetwork) -> MessageCache:
    """One operator-shaped message per directed edge, axes (bra link, ket link) (messagecache.jl:205-225)."""
    out = {}
    for v in nn.vertices():
        for w in nn.graph.neighbors(v):
            e = NamedEdge(w, v)
            ket = nn.ket.linkind(e)
            bra = Index(ket.dim, nn.braname(ket.name))
            out[e] = ITensor(np.zeros((ket.dim, ket.dim), dtype=nn.dtype), (bra, ket))
    return MessageCache(out)


def message_environment(f: Callable, nn: NormNetwork) -> MessageCache:
    return similar_message_environment(nn).map(f)


def ones_message(m: ITensor) -> ITensor:
    """`msg -> state(fill!(msg, true))` of test/test_apply_operator.jl:72."""
    return ITensor(np.ones_like(m.data), m.inds)


def identity_message(m: ITensor) -> ITensor:
    """`one` of test/test_apply_operator.jl:37."""
    return ITensor(np.eye(m.data.shape[0], dtype=m.data.dtype), m.inds)


# ---------------------------------------------------------------------------------------------------
# device session: factors + messages resident on one GPU
# ---------------------------------------------------------------------------------------------------
class _Session:
    def __init__(self, factors, device: int = 0, kernel: int = _lib.BPX_KERNEL_AUTO):
        self.factors = factors
        self.cp: CanonicalProblem = canonical_arrays(factors)
        ga = self.cp.ga
        self.ctx = BPXContext(device)
        self.ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        if kernel != _lib.BPX_KERNEL_AUTO:
            self.ctx.set_kernel_policy(kernel)
        self.ctx.set_dims(self.cp.dtype, self.cp.mode, self.cp.phys_dim if self.cp.mode == "norm" else None,
                          self.cp.link_dim)
        self.ctx.set_site_tensors(self.cp.tensors)
        self.version = 0  # bumped whenever the device message set changes (uploads, sweeps, single updates)

    def touched(self):
        self.version += 1

    def _msg_array(self, e: int, m) -> np.ndarray:
        if isinstance(m, ITensor):
            if self.cp.mode == "norm":
                return m.array(self.cp.bra_names[e], self.cp.ket_names[e])  # canonicalise to [bra, ket]
            return m.array(self.cp.ket_names[e])
        return np.asarray(m)

    def upload_messages(self, cache: MessageCache):
        ga = self.cp.ga
        msgs = []
        for e in range(ga.ne):
            ne_ = ga.named_edge(e)
            if ne_ not in cache:
                raise KeyError(f"no message on edge {ne_!r}")
            msgs.append(self._msg_array(e, cache[ne_]))
        self.ctx.set_messages(msgs)
        self.touched()
        cache._session_version = self.version

    def download_messages(self, like: MessageCache) -> MessageCache:
        ga = self.cp.ga
        arrs = self.ctx.get_messages()
        out = {}
        for e in range(ga.ne):
            ne_ = ga.named_edge(e)
            old = like[ne_] if ne_ in like else None
            if self.cp.mode == "norm":
                if isinstance(old, ITensor):
                    by = {i.name: i for i in old.inds}
                    inds = (by[self.cp.bra_names[e]], by[self.cp.ket_names[e]])
                else:
                    chi = self.cp.link_dim[e]
                    inds = (Index(chi, self.cp.bra_names[e]), Index(chi, self.cp.ket_names[e]))
            else:
                inds = old.inds if isinstance(old, ITensor) else (Index(self.cp.link_dim[e], self.cp.ket_names[e]),)
            out[ne_] = ITensor(arrs[e], inds)
        c = MessageCache(out)
        c._session = self
        c._session_version = self.version
        return c


def _session_for(factors, cache: MessageCache, device: int = 0) -> _Session:
    s = cache._session
    if s is not None and s.factors is factors:
        # the session may have moved on (another cache swept or updated it): make the device set equal THIS cache again
        if cache._session_version != s.version:
            s.upload_messages(cache)
        return s
    s = _Session(factors, device)
    s.upload_messages(cache)
    cache._session = s
    return s


# ---------------------------------------------------------------------------------------------------
# beliefpropagation (beliefpropagation.jl:69-92)
# ---------------------------------------------------------------------------------------------------
def default_beliefpropagation_edges(factors) -> List[NamedEdge]:
    return forest_cover_edge_sequence(factors.graph)


@dataclass
class BeliefPropagationResult:
    """What `StopWhenConvergedState` records (AIE.jl:67-71) plus the per-sweep residual history."""

    iterations: int = 0
    delta: float = math.inf
    at_iteration: int = -1
    residual_history: List[float] = field(default_factory=list)


def beliefpropagation(factors, messages, *, edges=None, stopping_criterion=None, message_update_algorithm=None,
                      info: Optional[BeliefPropagationResult] = None) -> MessageCache:
    """Run belief propagation on `factors` (a NormNetwork or a single-layer ITensorNetwork) starting
    from `messages` (dict / MessageCache keyed by directed edges).  Returns the final MessageCache."""
    cache = messages if isinstance(messages, MessageCache) else MessageCache(messages)
    alg = select_algorithm(message_update, message_update_algorithm, (cache, factors, None))
    criterion = select_beliefpropagation_stopping_criterion(stopping_criterion)
    if not isinstance(alg, (SimpleMessageUpdate, B200MessageUpdate)):
        raise TypeError(f"unsupported message update algorithm {type(alg).__name__}")
    try:
        maxiter, tol = _flatten_criterion(criterion)
    except _NotFlat:
        # a user-defined criterion: the reference's own construction (beliefpropagation.jl:75-91) over a device iterate,
        # one C-ABI call per sweep, the criterion evaluated on the host before every sweep
        if alg.schedule == "synchronous" and edges is not None:
            raise ArgumentError("`edges` selects the sequential schedule; the synchronous sweep updates every edge")
        if alg.schedule == "sequential" and edges is None:
            edges = default_beliefpropagation_edges(factors)
        n_steps = len(edges) if edges is not None else len(cache)
        sub = BeliefPropagationSweepAlgorithm(StopAfterIteration(n_steps), alg)
        algorithm = BeliefPropagationAlgorithm(edges, sub, criterion)
        session = _Session(factors, alg.device, getattr(alg, "kernel", _lib.BPX_KERNEL_AUTO))
        session.upload_messages(cache)
        return AI.solve(BeliefPropagationProblem(factors), algorithm, iterate=DeviceMessageCache(session, cache))
    if maxiter is None:
        maxiter = 2 ** 31 - 1
    session = _Session(factors, alg.device, getattr(alg, "kernel", _lib.BPX_KERNEL_AUTO))
    session.upload_messages(cache)
    if alg.schedule == "synchronous":
        if edges is not None:
            raise ArgumentError("`edges` selects the sequential schedule; the synchronous sweep updates every edge")
        res, done = session.ctx.sweep(maxiter, tol if tol is not None else 0.0, alg.normalize)
        session.touched()
    elif alg.schedule == "sequential":
        if edges is None:
            edges = default_beliefpropagation_edges(factors)
        seq = [session.cp.ga.edge_id(e) for e in edges]
        res, done = session.ctx.sweep_sequence(seq, maxiter, tol if tol is not None else 0.0, alg.normalize)
        session.touched()
    else:
        raise ArgumentError(f"unknown schedule {alg.schedule!r}")
    if info is not None:
        info.iterations = done
        info.delta = res
        info.residual_history = list(session.ctx.residual_history())
        info.at_iteration = done if (tol is not None and done > 0 and res < tol) else -1
    return session.download_messages(cache)


# ---------------------------------------------------------------------------------------------------
# The AlgorithmsInterface layer of BP (beliefpropagation.jl:94-210): problem / algorithm / state types, so that a
# caller can drive BP sweep by sweep with its own stopping criterion exactly as with the reference
# (`AI.solve(problem, algorithm; iterate = cache)`, `AI.step!`, `AI.is_finished!`).
# ---------------------------------------------------------------------------------------------------
class BeliefPropagationProblem(AI.Problem):
    def __init__(self, factors):
        self.factors = factors


class BeliefPropagationSweepProblem(AI.Problem):
    def __init__(self, factors, edges):
        self.factors = factors
        self.edges = edges


class BeliefPropagationSweepState(AI.State):
    def __init__(self, iterate, iteration: int = 0, stopping_criterion_state=None):
        self.iterate = iterate
        self.iteration = iteration
        self.stopping_criterion_state = stopping_criterion_state


class BeliefPropagationSweepAlgorithm(AI.Algorithm):
    """One sweep = `length(edges)` steps, each a `message_update!` of `edges[iteration]` (beliefpropagation.jl:160-210)."""

    def __init__(self, stopping_criterion: StoppingCriterion, message_update_algorithm=None):
        self.message_update_algorithm = SimpleMessageUpdate() if message_update_algorithm is None else message_update_algorithm
        self.stopping_criterion = stopping_criterion

    def initialize_state(self, problem, *, iterate, iteration: int = 0):
        scs = self.stopping_criterion.initialize_state(problem, self, iterate=iterate)
        return BeliefPropagationSweepState(iterate, iteration, scs)

    def step_(self, problem, state):
        edge = problem.edges[state.iteration - 1]  # Julia: problem.edges[state.iteration], 1-based
        message_update(state.iterate, problem.factors, edge, self.message_update_algorithm)
        return state


class BeliefPropagationState(AI.NestedState):
    def __init__(self, substate, iteration: int = 0, stopping_criterion_state=None):
        self.substate = substate
        self.iteration = iteration
        self.stopping_criterion_state = stopping_criterion_state


class BeliefPropagationAlgorithm(AI.NestedAlgorithm):
    """Outer loop: one step = one sweep (beliefpropagation.jl:100-151).

    The generic nested step runs the sweep edge by edge (one C-ABI call per `message_update!`).  When the iterate is a
    `DeviceMessageCache` the whole sweep is ONE call instead -- the Python twin of the `AI.step!` specialisation in
    julia/BPX.jl ("Why the dispatch hook sits one level above `message_update!`")."""

    def __init__(self, edges, subalgorithm: BeliefPropagationSweepAlgorithm, stopping_criterion: StoppingCriterion):
        self.edges = edges
        self.subalgorithm = subalgorithm
        self.stopping_criterion = stopping_criterion

    def initialize_state(self, problem, *, iterate, iteration: int = 0):
        subproblem = BeliefPropagationSweepProblem(problem.factors, self.edges)
        substate = self.subalgorithm.initialize_state(subproblem, iterate=iterate)
        scs = self.stopping_criterion.initialize_state(problem, self, iterate=iterate)
        return BeliefPropagationState(substate, iteration, scs)

    def initialize_subsolve(self, problem, state):
        return BeliefPropagationSweepProblem(problem.factors, self.edges), self.subalgorithm, state.substate

    def step_(self, problem, state):
        it = state.iterate
        if isinstance(it, DeviceMessageCache):
            it.sweep(self.edges, self.subalgorithm.message_update_algorithm)
            return state
        return super().step_(problem, state)

    def finalize_state_(self, problem, state):
        it = state.iterate
        return it.materialize() if isinstance(it, DeviceMessageCache) else it


class _DeviceSnapshot:
    """`copy(iterate)` of a device-resident iterate (StopWhenConverged copies it every outer iteration, AIE.jl:74, 96):
    a marker, not data -- the difference to the next iterate is the residual fused into the sweep kernels."""

    def __init__(self, owner: "DeviceMessageCache", version: int):
        self.owner = owner
        self.version = version

    def copy(self):
        return _DeviceSnapshot(self.owner, self.version)


class DeviceMessageCache(MessageCache):
    """A MessageCache whose messages live on the GPU between sweeps (SURVEY.md §8 b2 iii, variant B).

    `sweep` runs one sweep in one C-ABI call; `copy()` hands out a marker; `iterate_diff(previous)` returns the residual
    the update kernels fused into their epilogue (no second pass over the messages, no download).  Reading a message
    (`cache[edge]`, `items()`, ...) downloads the set once per version; `materialize()` gives an ordinary MessageCache."""

    def __init__(self, session: "_Session", like: MessageCache):
        self._session = session
        self._like = like
        self._version = 0
        self._host_version = -1
        self._host: Dict[NamedEdge, object] = {}
        self._last_residual = 0.0

    # the base class reads and writes `self._m`
    @property
    def _m(self):
        if self._host_version != self._version:
            self._host = dict(self._session.download_messages(self._like)._m)
            self._host_version = self._version
        return self._host

    @_m.setter
    def _m(self, value):
        self._host = value

    def __setitem__(self, e, m):
        raise ArgumentError("a DeviceMessageCache is updated by sweeps; use materialize() for a host copy to edit")

    def sweep(self, edges, alg) -> float:
        ctx, ga = self._session.ctx, self._session.cp.ga
        schedule = getattr(alg, "schedule", "sequential")
        if schedule == "synchronous":
            res, _ = ctx.sweep(1, 0.0, alg.normalize)
        else:
            seq = [ga.edge_id(e) for e in (edges if edges is not None else default_beliefpropagation_edges(self._session.factors))]
            res, _ = ctx.sweep_sequence(seq, 1, 0.0, alg.normalize)
        self._session.touched()
        self._version += 1
        self._last_residual = res
        return res

    def copy(self):
        return _DeviceSnapshot(self, self._version)

    def iterate_diff(self, other) -> float:
        if isinstance(other, _DeviceSnapshot) and other.owner is self:
            if other.version == self._version:
                return 0.0
            if other.version == self._version - 1:
                return self._last_residual
            raise ArgumentError("only the residual of the LAST sweep is kept on the device")
        ga = self._session.cp.ga
        return self._session.ctx.iterate_diff([self._session._msg_array(e, other[ga.named_edge(e)]) for e in range(ga.ne)])

    def materialize(self) -> MessageCache:
        c = self._session.download_messages(self._like)
        return c


def device_iterate(factors, messages, message_update_algorithm=None) -> DeviceMessageCache:
    """Upload factors and messages once and return the device-resident iterate to pass as `AI.solve(...; iterate=)`."""
    cache = messages if isinstance(messages, MessageCache) else MessageCache(messages)
    alg = select_algorithm(message_update, message_update_algorithm, (cache, factors, None))
    session = _Session(factors, alg.device, getattr(alg, "kernel", _lib.BPX_KERNEL_AUTO))
    session.upload_messages(cache)
    return DeviceMessageCache(session, cache)


def message_update(cache: MessageCache, factors, edge, alg=None, **kwargs) -> MessageCache:
    """Single in-place update `message_update!(cache, factors, edge)` (beliefpropagation.jl:230-257)."""
    alg = select_algorithm(message_update, alg, (cache, factors, edge), **kwargs)
    s = _session_for(factors, cache, alg.device)
    e = to_edge(edge)
    s.ctx.sweep_sequence([s.cp.ga.edge_id(e)], 1, 0.0, alg.normalize)
    s.touched()
    new = s.download_messages(cache)
    cache._m[e] = new[e]
    cache._session = s
    cache._session_version = s.version
    return cache


def iterate_diff(cache1: MessageCache, cache2: MessageCache) -> float:
    """max_e 1 - |<m1^, m2^>|^2 (beliefpropagation.jl:261-267), evaluated on the device."""
    s = cache1._session
    if s is None:
        raise ArgumentError("iterate_diff needs a cache produced by (or uploaded for) a device session")
    if cache1._session_version != s.version:  # the session has moved on: the device set must equal cache1
        s.upload_messages(cache1)
    ga = s.cp.ga
    other = [s._msg_array(e, cache2[ga.named_edge(e)]) for e in range(ga.ne)]
    return s.ctx.iterate_diff(other)


# ---------------------------------------------------------------------------------------------------
# beliefs (messagecache.jl:139-201)
# ---------------------------------------------------------------------------------------------------
def vertex_scalars(factors, messages: MessageCache, vertices=None):
    s = _session_for(factors, messages)
    vals = s.ctx.vertex_scalars()
    if vertices is None:
        return list(vals)
    return [vals[s.cp.ga.vindex[v]] for v in vertices]


def vertex_scalar(factors, messages: MessageCache, vertex):
    return vertex_scalars(factors, messages, [vertex])[0]


def _edge_session(messages: MessageCache, factors=None) -> _Session:
    if factors is not None:
        return _session_for(factors, messages)
    if messages._session is None:
        raise ArgumentError("edge scalars run on the device: pass `factors=` or a cache returned by beliefpropagation")
    if messages._session_version != messages._session.version:
        messages._session.upload_messages(messages)
    return messages._session


def edge_scalars(messages: MessageCache, factors=None):
    """One scalar per undirected edge (repeated edges and reverses ignored, messagecache.jl:161-178)."""
    return list(_edge_session(messages, factors).ctx.edge_scalars())


def edge_scalar(messages: MessageCache, edge, factors=None):
    s = _edge_session(messages, factors)
    ga = s.cp.ga
    e = ga.edge_id(edge)
    k = min(e, ga.rev[e])
    order = [x for x in range(ga.ne) if x < ga.rev[x]]
    return s.ctx.edge_scalars()[order.index(k)]


def region_scalar(factors, messages: MessageCache, region):
    out = 1.0
    for x in vertex_scalars(factors, messages, list(region)):
        out = out * x
    return out


def bethe_free_energy(factors, messages: MessageCache):
    """sum(log.(vertex scalars)) - sum(log.(edge scalars)) with the reference's complex promotion and -Inf rules
    (messagecache.jl:185-201), reduced on the device (`bpx_bethe_free_energy`): only the result crosses the bus."""
    s = _session_for(factors, messages)
    return s.ctx.bethe_free_energy()


def expect(factors: NormNetwork, messages: MessageCache, op: np.ndarray, vertices=None):
    """Local expectation values <O_v> = (vertex contraction with O on the ket site leg) / vertex_scalar.
    Build-defined extension (the reference has no `expect`); BASELINE.json's parity target
    names converged local expectation values."""
    s = _session_for(factors, messages)
    ops = [cast_to(op, s.cp.dtype, "operator")] * s.cp.ga.nv
    num = s.ctx.vertex_expect_numerators(ops)
    den = s.ctx.vertex_scalars()
    vals = num / den
    if vertices is None:
        return list(vals)
    return [vals[s.cp.ga.vindex[v]] for v in vertices]
