"""Host-side mirror of the iteration protocol the reference's BP driver is written against.

  AI   -- AlgorithmsInterface.jl 0.1 (external, not vendored; Project.toml:28).  Its observable contract is restated from
          the reference's call sites and tests (SURVEY.md §3.1 caveat): `Problem` / `Algorithm` / `State`,
          `initialize_state(!)`, `is_finished!` evaluated BEFORE every step, `increment!` before `step!` (1-based
          `problem.edges[state.iteration]`, beliefpropagation.jl:205), `solve` / `solve!`, `finalize_state!`
          (default: the iterate, beliefpropagation.jl:91 "-> typeof(cache)"; overridden at apply_operators.jl:123-127),
          `StopAfterIteration`, criteria combined with `|`   (test/test_algorithmsinterfaceextensions.jl:21-56, 133-138).
  AIE  -- src/AlgorithmsInterfaceExtensions/AlgorithmsInterfaceExtensions.jl: `NestedAlgorithm` (:7-32), `NestedState`
          (:39-53), `iterate_diff` (:59-61), `StopWhenConverged` (:63-119).

Julia's multiple dispatch becomes methods on the algorithm / criterion classes; the module-level functions keep the
reference's call shapes (`AI.solve(problem, algorithm; iterate = cache)`).  Functions ending in `!` end in `_` here.
Pure host logic: nothing in this file touches the device or the oracle.
"""
from __future__ import annotations

import math
from typing import Any, List, Optional


class MethodError(TypeError):
    """Julia `MethodError` analogue: no method of a generic function for these argument types."""


# =====================================================================================================================
# AI
# =====================================================================================================================
class Problem:
    pass


class StoppingCriterionState:
    pass


class StoppingCriterion:
    def __or__(self, other: "StoppingCriterion") -> "StopWhenAny":
        return StopWhenAny([self, other])

    # AI.initialize_state(problem, algorithm, criterion; iterate) -> criterion state
    def initialize_state(self, problem, algorithm, *, iterate=None) -> StoppingCriterionState:
        raise MethodError(f"initialize_state is not implemented for {type(self).__name__}")

    def initialize_state_(self, problem, algorithm, st: StoppingCriterionState) -> StoppingCriterionState:
        return st

    def is_finished(self, problem, algorithm, state, st) -> bool:
        raise MethodError(f"is_finished is not implemented for {type(self).__name__}")

    def is_finished_(self, problem, algorithm, state, st) -> bool:
        """Mutating check: records where the criterion fired."""
        if self.is_finished(problem, algorithm, state, st):
            st.at_iteration = state.iteration
            return True
        return False


class Algorithm:
    """Subclasses carry a `stopping_criterion` attribute and implement `initialize_state` and `step_`."""

    stopping_criterion: StoppingCriterion

    def initialize_state(self, problem, *, iterate, **kwargs) -> "State":
        raise MethodError(f"initialize_state is not implemented for {type(self).__name__}")

    def initialize_state_(self, problem, state: "State", *, iteration: int = 0, **kwargs) -> "State":
        for k, v in kwargs.items():
            setattr(state, k, v)
        state.iteration = iteration
        self.stopping_criterion.initialize_state_(problem, self, state.stopping_criterion_state)
        return state

    def step_(self, problem, state: "State") -> "State":
        raise MethodError(f"step! is not implemented for {type(self).__name__}")

    def increment_(self, problem, state: "State") -> "State":
        return increment_(state)

    def finalize_state_(self, problem, state: "State"):
        return state.iterate


class State:
    """Fields by convention: `iterate`, `iteration`, `stopping_criterion_state`."""

    iterate: Any
    iteration: int
    stopping_criterion_state: StoppingCriterionState


class _AtIterationState(StoppingCriterionState):
    def __init__(self):
        self.at_iteration = -1


class StopAfterIteration(StoppingCriterion):
    def __init__(self, maxiter: int):
        self.maxiter = int(maxiter)

    def __repr__(self):
        return f"StopAfterIteration({self.maxiter})"

    def __eq__(self, other):
        return isinstance(other, StopAfterIteration) and other.maxiter == self.maxiter

    def initialize_state(self, problem, algorithm, *, iterate=None):
        return _AtIterationState()

    def initialize_state_(self, problem, algorithm, st):
        st.at_iteration = -1
        return st

    def is_finished(self, problem, algorithm, state, st) -> bool:
        return state.iteration >= self.maxiter


class _AnyState(StoppingCriterionState):
    def __init__(self, states: List[StoppingCriterionState]):
        self.states = states
        self.at_iteration = -1


class StopWhenAny(StoppingCriterion):
    """`c1 | c2`: finished when any member is.  Every member is evaluated each time (no short circuit), so stateful
    members (StopWhenConverged's previous iterate) are refreshed every iteration."""

    def __init__(self, criteria: List[StoppingCriterion]):
        self.criteria = list(criteria)

    def __or__(self, other):
        return StopWhenAny(self.criteria + [other])

    def __repr__(self):
        return " | ".join(repr(c) for c in self.criteria)

    def __eq__(self, other):
        return isinstance(other, StopWhenAny) and other.criteria == self.criteria

    def initialize_state(self, problem, algorithm, *, iterate=None):
        return _AnyState([c.initialize_state(problem, algorithm, iterate=iterate) for c in self.criteria])

    def initialize_state_(self, problem, algorithm, st):
        for c, s in zip(self.criteria, st.states):
            c.initialize_state_(problem, algorithm, s)
        st.at_iteration = -1
        return st

    def is_finished(self, problem, algorithm, state, st) -> bool:
        return any(c.is_finished(problem, algorithm, state, s) for c, s in zip(self.criteria, st.states))

    def is_finished_(self, problem, algorithm, state, st) -> bool:
        hits = [c.is_finished_(problem, algorithm, state, s) for c, s in zip(self.criteria, st.states)]
        if any(hits):
            st.at_iteration = state.iteration
            return True
        return False


def initialize_state(problem, algorithm, criterion: Optional[StoppingCriterion] = None, *, iterate=None, **kwargs):
    if criterion is not None:
        return criterion.initialize_state(problem, algorithm, iterate=iterate)
    return algorithm.initialize_state(problem, iterate=iterate, **kwargs)


def initialize_state_(problem, algorithm, target, st: Optional[StoppingCriterionState] = None, **kwargs):
    """`AI.initialize_state!(problem, algorithm, state; kw...)` or `(problem, algorithm, criterion, criterion_state)`."""
    if isinstance(target, StoppingCriterion):
        return target.initialize_state_(problem, algorithm, st)
    return algorithm.initialize_state_(problem, target, **kwargs)


def increment_(state: State) -> State:
    state.iteration += 1
    return state


def is_finished_(problem, algorithm, state: State) -> bool:
    return algorithm.stopping_criterion.is_finished_(problem, algorithm, state, state.stopping_criterion_state)


def is_finished(problem, algorithm, state: State) -> bool:
    return algorithm.stopping_criterion.is_finished(problem, algorithm, state, state.stopping_criterion_state)


def step_(problem, algorithm, state: State) -> State:
    return algorithm.step_(problem, state)


def solve_(problem, algorithm, state: State, **kwargs):
    """`AI.solve!`: re-initialise, then  while !is_finished!: increment!, step!  and finalize."""
    algorithm.initialize_state_(problem, state, **kwargs)
    while not is_finished_(problem, algorithm, state):
        algorithm.increment_(problem, state)
        algorithm.step_(problem, state)
    return algorithm.finalize_state_(problem, state)


def solve(problem, algorithm, **kwargs):
    """`AI.solve(problem, algorithm; iterate, ...)`: fresh state, then `solve!`."""
    state = algorithm.initialize_state(problem, **kwargs)
    return solve_(problem, algorithm, state)


# =====================================================================================================================
# AIE
# =====================================================================================================================
class NestedAlgorithm(Algorithm):
    """One outer step = one inner `solve!` (AIE.jl:7-32).  Subclasses override `initialize_subsolve`."""

    def initialize_subsolve(self, problem, state):
        raise MethodError(f"initialize_subsolve is not implemented for {type(self).__name__}")

    def finalize_substate_(self, problem, state, substate):
        state.iterate = substate.iterate
        return state

    def step_(self, problem, state):
        subproblem, subalgorithm, substate = self.initialize_subsolve(problem, state)
        solve_(subproblem, subalgorithm, substate)
        self.finalize_substate_(problem, state, substate)
        return state


def initialize_subsolve(problem, algorithm, state):
    """Generic default (AIE.jl:14-18): a MethodError unless the algorithm provides its own."""
    f = getattr(algorithm, "initialize_subsolve", None)
    if f is None:
        raise MethodError(f"initialize_subsolve({type(problem).__name__}, {type(algorithm).__name__}, {type(state).__name__})")
    return f(problem, state)


def finalize_substate_(problem, algorithm, state, substate):
    """AIE.jl:20-25: copy the substate's iterate back into the parent state."""
    f = getattr(algorithm, "finalize_substate_", None)
    if f is not None:
        return f(problem, state, substate)
    state.iterate = substate.iterate
    return state


class NestedState(State):
    """Forwards `iterate` to `self.substate.iterate` (AIE.jl:39-53)."""

    substate: State

    @property
    def iterate(self):
        return self.substate.iterate

    @iterate.setter
    def iterate(self, value):
        self.substate.iterate = value


def iterate_diff(a, b) -> float:
    """AIE.jl:59-61: concrete iterate types supply the method (here: an `iterate_diff` method on the iterate)."""
    f = getattr(a, "iterate_diff", None)
    if f is None:
        raise MethodError(f"iterate_diff({type(a).__name__}, {type(b).__name__})")
    return f(b)


class StopWhenConvergedState(StoppingCriterionState):
    def __init__(self, previous_iterate):
        self.delta = math.inf
        self.at_iteration = -1
        self.previous_iterate = previous_iterate


class StopWhenConverged(StoppingCriterion):
    """Fires once `iterate_diff(iterate, previous_iterate) < tol` (AIE.jl:63-119)."""

    def __init__(self, tol: float):
        self.tol = float(tol)  # `tol::Float64`

    def __repr__(self):
        return f"StopWhenConverged({self.tol})"

    def __eq__(self, other):
        return isinstance(other, StopWhenConverged) and other.tol == self.tol

    def initialize_state(self, problem, algorithm, *, iterate=None):
        return StopWhenConvergedState(previous_iterate=_copy(iterate))

    def initialize_state_(self, problem, algorithm, st):
        st.delta = math.inf
        return st

    def is_finished_(self, problem, algorithm, state, st) -> bool:
        iterate = state.iterate
        delta = iterate_diff(iterate, st.previous_iterate)
        st.previous_iterate = _copy(iterate)
        if state.iteration == 0:  # delta = 0 initially, so skip this the first time
            return False
        st.delta = delta
        if self.is_finished(problem, algorithm, state, st):
            st.at_iteration = state.iteration
            return True
        return False

    def is_finished(self, problem, algorithm, state, st) -> bool:
        return st.delta < self.tol


def _copy(x):
    return x.copy() if hasattr(x, "copy") else list(x)
