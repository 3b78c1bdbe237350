"""The iteration protocol mirror (itensornetworksnext.jl_b200/algorithmsinterface.py) against the reference's own tests
(test/test_algorithmsinterfaceextensions.jl:107-139) and the contract they document (:21-56), plus the AI layer of BP
(beliefpropagation.jl:94-210) driven the way a reference user drives it: `AI.solve`, manual `AI.step!` / `AI.is_finished!`,
user-defined stopping criteria.  Device work goes through tests/native_ctx.HostHarnessContext on the CPU (BP sweeps from
the numpy oracle -- test infrastructure) and through libbpx.so in the `gpu` variants (tests/test_zzz_late_gpu.py)."""
import math
import sys

import numpy as np
import pytest

import itnn_b200 as B
from itnn_b200 import AI, AIE
from itnn_b200 import graphs
from native_ctx import HostHarnessContext

bp_mod = sys.modules["itnn_b200.beliefpropagation"]  # the package attribute of that name is the function


# ---- the reference's test fixtures (test/test_algorithmsinterfaceextensions.jl:9-104) --------------------------------
class TestProblem(AI.Problem):
    __test__ = False


class TestChildState(AI.State):
    __test__ = False

    def __init__(self, iterate, stopping_criterion_state, iteration=0):
        self.iterate = iterate
        self.iteration = iteration
        self.stopping_criterion_state = stopping_criterion_state


class TestChildAlgorithm(AI.Algorithm):
    __test__ = False

    def __init__(self, stopping_criterion=None):
        self.stopping_criterion = AI.StopAfterIteration(2) if stopping_criterion is None else stopping_criterion

    def initialize_state(self, problem, *, iterate, **kwargs):
        sc_state = AI.initialize_state(problem, self, self.stopping_criterion, iterate=iterate)
        return TestChildState(iterate, sc_state, **kwargs)

    def step_(self, problem, state):
        state.iterate += 1  # `state.iterate .+= 1`: in place
        return state


class TestNestedAlgorithm(AIE.NestedAlgorithm):
    __test__ = False

    def __init__(self, algorithms, stopping_criterion=None):
        self.algorithms = algorithms
        self.stopping_criterion = AI.StopAfterIteration(len(algorithms)) if stopping_criterion is None else stopping_criterion

    def initialize_state(self, problem, *, iterate, **kwargs):
        sc_state = AI.initialize_state(problem, self, self.stopping_criterion, iterate=iterate)
        return TestChildState(iterate, sc_state, **kwargs)

    def initialize_subsolve(self, problem, state):
        subalgorithm = self.algorithms[state.iteration - 1]
        substate = AI.initialize_state(problem, subalgorithm, iterate=state.iterate)
        return problem, subalgorithm, substate


def test_nested_algorithm_defaults():  # :108-121
    problem, algorithm = TestProblem(), TestChildAlgorithm()
    state = AI.initialize_state(problem, algorithm, iterate=np.array([0.0]))
    with pytest.raises(B.MethodError):
        AIE.initialize_subsolve(problem, algorithm, state)
    with pytest.raises(TypeError):  # MethodError is the TypeError the rest of the mirror raises for missing methods
        AIE.NestedAlgorithm().initialize_subsolve(problem, state)
    substate = AI.initialize_state(problem, algorithm, iterate=np.array([42.0]))
    AIE.finalize_substate_(problem, algorithm, state, substate)
    assert np.array_equal(state.iterate, [42.0])


def test_nested_algorithm_runs_the_children():  # :123-139
    problem = TestProblem()
    nested = TestNestedAlgorithm([TestChildAlgorithm(AI.StopAfterIteration(1)), TestChildAlgorithm(AI.StopAfterIteration(2))])
    assert isinstance(nested, AIE.NestedAlgorithm)
    state = AI.initialize_state(problem, nested, iterate=np.array([0.0, 0.0]))
    AI.solve_(problem, nested, state, iterate=np.array([0.0, 0.0]))
    assert state.iteration == 2                     # two child algorithms ...
    assert np.allclose(state.iterate, [3.0, 3.0])   # ... 1 + 2 inner steps


def test_solve_protocol_and_criteria():
    """is_finished! before every step, increment! before step!, `|` of criteria, at_iteration bookkeeping."""
    problem = TestProblem()
    alg = TestChildAlgorithm(AI.StopAfterIteration(5))
    out = AI.solve(problem, alg, iterate=np.array([0.0]))
    assert np.array_equal(out, [5.0])               # default finalize_state!: the iterate
    assert AI.StopAfterIteration(3) | AI.StopAfterIteration(7) == AI.StopWhenAny([AI.StopAfterIteration(3), AI.StopAfterIteration(7)])
    alg = TestChildAlgorithm(AI.StopAfterIteration(7) | AI.StopAfterIteration(3))
    state = AI.initialize_state(problem, alg, iterate=np.array([0.0]))
    AI.solve_(problem, alg, state)
    assert state.iteration == 3 and state.stopping_criterion_state.at_iteration == 3
    assert [s.at_iteration for s in state.stopping_criterion_state.states] == [-1, 3]
    AI.solve_(problem, alg, state, iterate=np.array([10.0]))  # solve! re-initialises: iteration 0, criterion states reset
    assert state.iteration == 3 and np.array_equal(state.iterate, [13.0])


class Vec:
    """An iterate with its own `iterate_diff` (what AIE.jl:59-61 asks of concrete iterate types)."""

    def __init__(self, x):
        self.x = np.asarray(x, dtype=float)

    def copy(self):
        return Vec(self.x.copy())

    def iterate_diff(self, other):
        return float(np.abs(self.x - other.x).max())


class Halve(AI.Algorithm):
    def __init__(self, stopping_criterion):
        self.stopping_criterion = stopping_criterion

    def initialize_state(self, problem, *, iterate, **kwargs):
        return TestChildState(iterate, AI.initialize_state(problem, self, self.stopping_criterion, iterate=iterate), **kwargs)

    def step_(self, problem, state):
        state.iterate.x *= 0.5
        return state


def test_stop_when_converged():  # AIE.jl:63-119
    problem = TestProblem()
    alg = Halve(AI.StopAfterIteration(100) | AIE.StopWhenConverged(1e-3))
    state = AI.initialize_state(problem, alg, iterate=Vec([1.0]))
    conv = state.stopping_criterion_state.states[1]
    assert conv.delta == math.inf and conv.at_iteration == -1
    AI.solve_(problem, alg, state)
    # x_k = 2^-k, delta_k = 2^-k: the first delta < 1e-3 is k = 10; iteration 0 never stops (delta = 0 there)
    assert state.iteration == 10 and conv.at_iteration == 10 and conv.delta == 2.0 ** -10
    with pytest.raises(B.MethodError):
        AIE.iterate_diff([1.0], [1.0])  # no method for plain lists
    assert AIE.StopWhenConverged(1).tol == 1.0 and isinstance(AIE.StopWhenConverged(1).tol, float)


# ---- the AI layer of BP ----------------------------------------------------------------------------------------------
@pytest.fixture
def host_ctx(monkeypatch):
    from native_ctx import build_hostlib

    lib = build_hostlib()
    monkeypatch.setattr(bp_mod, "BPXContext", lambda device=0: HostHarnessContext(lib, device))


def peps(dtype=np.float64):
    g = graphs.named_grid((3, 3))
    tn, _, _ = B.random_state(dtype, g, d=2, chi=2, rng=np.random.default_rng(4))
    nn = B.normnetwork(tn)
    return nn, B.message_environment(B.identity_message, nn)


class StopWhenSlow(AI.StoppingCriterion):
    """A user-defined criterion: stop when the residual shrinks by less than a factor 2 per sweep, or after 30 sweeps."""

    class St(AI.StoppingCriterionState):
        def __init__(self, previous):
            self.previous, self.last, self.at_iteration, self.deltas = previous, math.inf, -1, []

    def initialize_state(self, problem, algorithm, *, iterate=None):
        return self.St(iterate.copy())

    def is_finished_(self, problem, algorithm, state, st):
        delta = AIE.iterate_diff(state.iterate, st.previous)
        st.previous = state.iterate.copy()
        if state.iteration == 0:
            return False
        st.deltas.append(delta)
        slow = delta > 0.5 * st.last or state.iteration >= 30
        st.last = delta
        if slow:
            st.at_iteration = state.iteration
        return slow


def check_bp_ai_layer(schedule):
    nn, env0 = peps()
    alg = B.B200MessageUpdate(schedule=schedule)
    crit = dict(maxiter=6, tol=1e-30)
    want = B.beliefpropagation(nn, env0, message_update_algorithm=alg, stopping_criterion=crit)  # fused fast path
    edges = None if schedule == "synchronous" else B.default_beliefpropagation_edges(nn)
    n_steps = len(env0) if edges is None else len(edges)

    # 1. the reference's own construction (beliefpropagation.jl:75-91) over a device iterate, AI.solve
    problem = B.BeliefPropagationProblem(nn)
    sub = B.BeliefPropagationSweepAlgorithm(AI.StopAfterIteration(n_steps), alg)
    algorithm = B.BeliefPropagationAlgorithm(edges, sub, B.select_beliefpropagation_stopping_criterion(crit))
    got = AI.solve(problem, algorithm, iterate=B.device_iterate(nn, env0, alg))
    assert type(got) is B.MessageCache
    assert all(np.allclose(got[e].data, want[e].data, rtol=1e-12, atol=1e-15) for e in want.keys())

    # 2. stepping by hand: is_finished! / increment! / step!, the fused residual read through StopWhenConverged's state
    state = AI.initialize_state(problem, algorithm, iterate=B.device_iterate(nn, env0, alg))
    deltas = []
    while not AI.is_finished_(problem, algorithm, state):
        AI.increment_(state)
        AI.step_(problem, algorithm, state)
        deltas.append(state.iterate._last_residual)
    conv = state.stopping_criterion_state.states[1]
    assert state.iteration == 6 and conv.delta == deltas[-1] and deltas[-1] < deltas[0]
    assert isinstance(state, AIE.NestedState) and state.iterate is state.substate.iterate
    # the device residual IS iterate_diff of consecutive iterates (checked against materialised copies)
    before = state.iterate.materialize()
    AI.step_(problem, algorithm, state)
    assert abs(state.iterate._last_residual - B.iterate_diff(state.iterate.materialize(), before)) < 1e-13

    # 3. a user-defined criterion through `beliefpropagation` itself
    mine = StopWhenSlow()
    out = B.beliefpropagation(nn, env0, message_update_algorithm=alg, stopping_criterion=mine)
    assert type(out) is B.MessageCache and len(out) == len(env0)

    # 4. the literal per-edge path: a plain MessageCache iterate -> one message_update! per step (sequential only)
    if schedule == "sequential":
        lit = B.BeliefPropagationAlgorithm(edges, B.BeliefPropagationSweepAlgorithm(AI.StopAfterIteration(n_steps), B.SimpleMessageUpdate()),
                                           AI.StopAfterIteration(1))
        one = AI.solve(problem, lit, iterate=env0.copy())
        ref = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=1))
        assert all(np.allclose(one[e].data, ref[e].data, rtol=1e-12, atol=1e-15) for e in ref.keys())


@pytest.mark.parametrize("schedule", ["synchronous", "sequential"])
def test_bp_ai_layer_host_harness(host_ctx, schedule):
    check_bp_ai_layer(schedule)


def test_device_message_cache_is_read_only_and_lazy(host_ctx):
    nn, env0 = peps()
    it = B.device_iterate(nn, env0, B.B200MessageUpdate())
    snap = it.copy()
    assert it.iterate_diff(snap) == 0.0
    it.sweep(None, B.B200MessageUpdate())
    assert it.iterate_diff(snap) == it._last_residual > 0
    it.sweep(None, B.B200MessageUpdate())
    with pytest.raises(B.ArgumentError, match="LAST sweep"):
        it.iterate_diff(snap)
    e = next(iter(env0.keys()))
    with pytest.raises(B.ArgumentError):
        it[e] = env0[e]
    assert it[e].data.shape == env0[e].data.shape and len(it) == len(env0)
    assert abs(it.iterate_diff(env0) - B.iterate_diff(it.materialize(), env0)) < 1e-13
