"""`ResidentState`: a state and its BP environment kept RESIDENT on one B200 across BP runs and gate layers.

The reference composes `beliefpropagation` (src/beliefpropagation/beliefpropagation.jl:69-92) and `apply_operators`
(src/apply/apply_operators.jl:28-60) on the host, copying state and environment on every call (`initialize_output`,
:204-208).  A simple-update evolution alternates the two thousands of times, so the B200-first shape of that loop is one
device context that owns the site tensors and the message set; gate layers (`bpx_apply_*_gates`), sweeps (`bpx_sweep`,
`bpx_sweep_sequence`) and local expectation values (`bpx_vertex_expect_numerators`) all run on it without a host round
trip.  Same arithmetic, same entry points as `apply.py` / `beliefpropagation.py`; link dimensions are fixed for the life
of the object (the fixed-chi simple update: ranks above the leg's dimension are truncated, lower ranks zero-padded).

No arithmetic of the path runs here; there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .device import cast_to

from .apply import (Operator, _ApplySession, _lowered_operator, _reference_rank, _touched_vertices)
from .beliefpropagation import (ArgumentError, BeliefPropagationResult, MessageCache, _flatten_criterion, _NotFlat,
                                select_beliefpropagation_stopping_criterion)
from .graphs import NamedEdge, forest_cover_edge_sequence
from .tensornetwork import Index, ITensor, ITensorNetwork


class ResidentState:
    def __init__(self, state: ITensorNetwork, env: Optional[MessageCache] = None, device: int = 0):
        """`env`: operator-shaped messages M_e[bra, ket] for every directed edge; None = all ones
        (`message_environment(ones_message, nn)`, test/test_apply_operator.jl:72)."""
        if env is not None and not isinstance(env, MessageCache):
            env = MessageCache(env)
        self._s = _ApplySession(state, env, device)
        self.names = state  # names and dimensions of the legs (tensor DATA lives on the device)

    # -- plumbing --------------------------------------------------------------------------------------------
    @property
    def ctx(self):
        return self._s.ctx

    def close(self):
        self._s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- gates (apply_operators.jl:213-283) ---------------------------------------------------------------------
    def apply_layer(self, operators: Sequence[Operator], trunc: Optional[int] = None, normalize: bool = False) -> List[np.ndarray]:
        """One device call for a set of vertex-disjoint gates of one kind.  Returns the kept singular values per
        two-site gate ([] for one-site gates)."""
        operators = list(operators)
        if not operators:
            return []
        s, cp = self._s, self._s.cp
        touched = [_touched_vertices(op, self.names) for op in operators]
        if len({len(vs) for vs in touched}) != 1:
            raise ArgumentError("a layer holds one-site gates or two-site gates, not both")
        flat = [v for vs in touched for v in vs]
        if len(set(flat)) != len(flat):
            raise ArgumentError("the gates of a layer must be vertex-disjoint")
        lowered = [_lowered_operator(op, self.names, vs, cp.dtype) for op, vs in zip(operators, touched)]
        if len(touched[0]) == 1:
            s.ctx.apply_one_site_gates([cp.ga.vindex[vs[0]] for vs in touched], lowered, normalize=normalize)
            return []
        edges = []
        for vs in touched:
            if vs[1] not in self.names.graph.neighbors(vs[0]):
                raise ArgumentError(f"two-site gate on vertices {vs[0]!r}, {vs[1]!r} that share no link")
            e = cp.ga.edge_id(NamedEdge(vs[0], vs[1]))
            if _reference_rank(cp, e, trunc) > s.link_dim[e]:
                raise ArgumentError(
                    f"the gate on {vs[0]!r}-{vs[1]!r} would grow its bond beyond {s.link_dim[e]}: a resident state keeps its "
                    "link dimensions; pass `trunc` <= the bond dimension (or start from a state padded to the target chi)")
            edges.append(e)
        return s.ctx.apply_two_site_gates(edges, lowered, max_rank=0 if trunc is None else int(trunc), normalize=normalize)

    def apply_operators(self, operators: Sequence[Operator], trunc: Optional[int] = None, normalize: bool = False):
        """Operators applied in turn (apply_operators.jl:106-121); consecutive vertex-disjoint gates of one kind share a
        device call (they commute, so the result is the sequential one)."""
        batch, used, kind = [], set(), None
        for op in operators:
            vs = _touched_vertices(op, self.names)
            if batch and (len(vs) != kind or used & set(vs)):
                self.apply_layer(batch, trunc, normalize)
                batch, used = [], set()
            batch.append(op)
            used |= set(vs)
            kind = len(vs)
        if batch:
            self.apply_layer(batch, trunc, normalize)
        return self

    # -- belief propagation on the norm network of the resident state (beliefpropagation.jl:69-92) -------------------
    def beliefpropagation(self, stopping_criterion=None, schedule: str = "synchronous", normalize: bool = True,
                          edges=None) -> BeliefPropagationResult:
        try:
            maxiter, tol = _flatten_criterion(select_beliefpropagation_stopping_criterion(stopping_criterion))
        except _NotFlat:
            raise ArgumentError("ResidentState.beliefpropagation takes StopAfterIteration / StopWhenConverged criteria "
                                "(maxiter, tol); drive custom criteria through AI.solve with a DeviceMessageCache") from None
        maxiter = 2 ** 31 - 1 if maxiter is None else maxiter
        ctx, ga = self._s.ctx, self._s.cp.ga
        if schedule == "synchronous":
            if edges is not None:
                raise ArgumentError("`edges` selects the sequential schedule; the synchronous sweep updates every edge")
            res, done = ctx.sweep(maxiter, tol if tol is not None else 0.0, normalize)
        elif schedule == "sequential":
            seq = [ga.edge_id(e) for e in (forest_cover_edge_sequence(self.names.graph) if edges is None else edges)]
            res, done = ctx.sweep_sequence(seq, maxiter, tol if tol is not None else 0.0, normalize)
        else:
            raise ArgumentError(f"unknown schedule {schedule!r}")
        return BeliefPropagationResult(iterations=done, delta=res, residual_history=list(ctx.residual_history()),
                                       at_iteration=done if (tol is not None and done > 0 and res < tol) else -1)

    # -- beliefs (messagecache.jl:139-201; `expect` is the build-defined extension) ---------------------
    def vertex_scalars(self):
        return list(self._s.ctx.vertex_scalars())

    def edge_scalars(self):
        return list(self._s.ctx.edge_scalars())

    def expect(self, op: np.ndarray, vertices=None):
        cp = self._s.cp
        ops = [cast_to(op, cp.dtype, "operator")] * cp.ga.nv
        vals = self._s.ctx.vertex_expect_numerators(ops) / self._s.ctx.vertex_scalars()
        if vertices is None:
            return list(vals)
        return [vals[cp.ga.vindex[v]] for v in vertices]

    def expect_two_site(self, operators: Sequence[Operator]):
        """<O> for two-site operators on neighbouring vertices in the BP environment (bond energies of a simple-update
        evolution); read-only, so the operators may overlap.  Build-defined like `expect`."""
        operators = list(operators)
        cp = self._s.cp
        edges, lowered = [], []
        for op in operators:
            vs = _touched_vertices(op, self.names)
            if len(vs) != 2 or vs[1] not in self.names.graph.neighbors(vs[0]):
                raise ArgumentError("expect_two_site takes operators on two neighbouring vertices")
            edges.append(cp.ga.edge_id(NamedEdge(vs[0], vs[1])))
            lowered.append(_lowered_operator(op, self.names, vs, cp.dtype))
        num, den = self._s.ctx.edge_expect(edges, lowered)
        return list(num / den)

    # -- back to the host ------------------------------------------------------------------------------------------
    def state(self) -> ITensorNetwork:
        return ITensorNetwork({v: self._s.site_itensor(v) for v in self.names.vertices()})

    def env(self) -> MessageCache:
        cp = self._s.cp
        out = {}
        for e, m in enumerate(self._s.ctx.get_messages()):
            ket = cp.ket_names[e]
            c = m.shape[0]
            out[cp.ga.named_edge(e)] = ITensor(m, (Index(c, cp.bra_names[e]), Index(c, ket)))
        return MessageCache(out)
