"""Single-layer (plain ITensorNetwork factors, vector messages) synchronous sweeps on the GPU: the thread-per-vertex
register kernel (csrc/bpx_vertex.cuh, BPX_KERNEL_VERTEX) and the generic kernel against the CPU oracle, the
ITensorNetworkGenerators inputs end to end, and the bench workload's large lattice.  SURVEY.md §8 f1.

(The file sorts last on purpose: these are the newest kernels, and `pytest -x` should reach every older test first.)"""
import math

import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn, rel_err, single_layer_tensors
from itnn_b200 import _lib, generators, graphs, problems
from test_generators import f_1d_ising_open, link_function
from test_gpu_parity import MSG_RTOL, check_sweeps, make_ctx

pytestmark = pytest.mark.gpu

VERTEX, GENERIC, AUTO = _lib.BPX_KERNEL_VERTEX, _lib.BPX_KERNEL_GENERIC, _lib.BPX_KERNEL_AUTO


def random_vectors(ga, link_dim, dtype, rng):
    out = []
    for e in range(ga.ne):
        m = rng.random(link_dim[e]) + 0.1
        if np.dtype(dtype).kind == "c":
            m = m + 0.3j * rng.standard_normal(link_dim[e])
        out.append((m / m.sum()).astype(dtype))
    return out


def vertex_kernel_takes(dtype, z, chi):
    w = 2 if np.dtype(dtype).kind == "c" else 1
    return 1 <= z <= 6 and 2 <= chi <= 4 and chi ** z * w <= 64


CASES = {
    # name: (graph, chi)
    "torus4x4_chi2": (lambda: graphs.named_grid((4, 4), periodic=True), 2),        # degree 4 everywhere (Ising / spin ice)
    "open5x4_chi2": (lambda: graphs.named_grid((5, 4)), 2),                        # degrees 2, 3, 4: three launches
    "cubic3_chi2": (lambda: graphs.named_grid((3, 3, 3), periodic=True), 2),       # degree 6: 64 values per thread
    "comb_chi3": (lambda: graphs.named_comb_tree((4, 3)), 3),                      # degrees 1, 2, 3; odd dimension
    "open4x4_chi4": (lambda: graphs.named_grid((4, 4)), 4),                        # degree 4 (256 values) stays generic
    "heavyhex_chi2": (graphs.heavy_hex_127, 2),                                    # irregular graph
    "chain7_chi4": (lambda: graphs.named_path_graph(7), 4),                        # degrees 1 and 2
    "open3x3_chi5": (lambda: graphs.named_grid((3, 3)), 5),                        # chi = 5: all generic
}


@pytest.mark.parametrize("kernel", [AUTO, GENERIC])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("case", sorted(CASES))
def test_single_layer_synchronous_sweeps(oracle, case, dtype, kernel):
    make_graph, chi = CASES[case]
    ga = graphs.graph_arrays(make_graph())
    rng = np.random.default_rng(5)
    tensors = single_layer_tensors(ga, chi, dtype, rng)
    msgs = random_vectors(ga, [chi] * ga.ne, dtype, rng)
    buckets = check_sweeps(oracle, ga, dtype, "single", None, [chi] * ga.ne, tensors, msgs, 3, kernel)
    for b in buckets:
        want = VERTEX if (kernel == AUTO and vertex_kernel_takes(dtype, b["degree"], b["chi"])) else GENERIC
        assert b["kernel"] == want, (b, want)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_vertex_kernel_mixed_link_dims_fall_back_per_bucket(oracle, dtype):
    # per-leg different dims are not a VERTEX shape: those buckets run on the generic kernel in the same sweep
    g = graphs.named_grid((4, 3))
    ga = graphs.graph_arrays(g)
    rng = np.random.default_rng(11)
    link_dim = [0] * ga.ne
    for e in range(ga.ne):
        if e < ga.rev[e]:
            link_dim[e] = link_dim[ga.rev[e]] = 2 if ga.src[e] % 4 else 3
    tensors = [randn(rng, dtype, tuple(link_dim[f] for f in range(ga.row_ptr[v], ga.row_ptr[v + 1]))) for v in range(ga.nv)]
    msgs = random_vectors(ga, link_dim, dtype, rng)
    buckets = check_sweeps(oracle, ga, dtype, "single", None, link_dim, tensors, msgs, 2)
    kinds = {b["kernel"] for b in buckets}
    assert kinds == {VERTEX, GENERIC}, buckets


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_vertex_kernel_zero_sum_guard_no_normalize_and_nan(oracle, dtype):
    g = graphs.named_path_graph(3)
    ga = graphs.graph_arrays(g)
    # the middle factor sends [1, -1] (sum exactly 0) to its second neighbour: left unnormalised (beliefpropagation.jl:248-253)
    t_mid = np.array([[1.0, -1.0], [0.0, 0.0]], dtype=dtype)
    tensors = [np.array([1.0, 2.0], dtype=dtype), t_mid, np.array([0.5, 0.25], dtype=dtype)]
    msgs = [np.array([1.0, 0.0], dtype=dtype) for _ in range(ga.ne)]
    p = oracle.make_problem(ga, tensors, "single")
    want = oracle.sweep_jacobi(p, msgs)
    assert any(abs(w.sum()) == 0 and np.abs(w).max() > 0 for w in want), "the case must exercise the iszero branch"
    with make_ctx(ga, dtype, "single", None, [2] * ga.ne, tensors, msgs) as ctx:
        assert all(b["kernel"] == VERTEX for b in ctx.buckets())
        ctx.sweep(1)
        got = ctx.get_messages()
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-15)
    # normalize = false
    rng = np.random.default_rng(2)
    g = graphs.named_grid((3, 3), periodic=True)
    ga = graphs.graph_arrays(g)
    tensors = single_layer_tensors(ga, 2, dtype, rng)
    msgs = random_vectors(ga, [2] * ga.ne, dtype, rng)
    check_sweeps(oracle, ga, dtype, "single", None, [2] * ga.ne, tensors, msgs, 2, normalize=False)
    # a NaN in one factor reaches the sweep's residual (Julia's `maximum` propagates NaN)
    tensors[4] = tensors[4].copy()
    tensors[4].flat[3] = np.nan
    with make_ctx(ga, dtype, "single", None, [2] * ga.ne, tensors, msgs) as ctx:
        res, _ = ctx.sweep(1)
        assert math.isnan(res)


def test_ising_generator_networks_end_to_end(oracle):
    """`ising_network` inputs through the reference-shaped API with the B200 strategy (synchronous sweeps)."""
    beta = 0.4
    alg = B.B200MessageUpdate()
    # open chain (a tree): BP is exact after diameter-many synchronous sweeps -> analytic free energy (test/utils.jl:28-39)
    n = 9
    g = graphs.named_grid((n,))
    tn = generators.ising_network(link_function(g), beta, g)
    ones = {e: B.ITensor(np.ones(2), (tn.linkind(e),)) for e in g.all_edges()}
    cache = B.beliefpropagation(tn, ones, stopping_criterion=dict(maxiter=n), message_update_algorithm=alg)
    f_bp = -B.bethe_free_energy(tn, cache) / (beta * n)
    assert np.isclose(f_bp, f_1d_ising_open(beta, n), rtol=1e-12)
    # 4x4 torus: loopy; the converged Bethe free energy equals the oracle's
    g = graphs.named_grid((4, 4), periodic=True)
    tn = generators.ising_network(link_function(g), beta, g)
    rng = np.random.default_rng(123)
    init = {e: B.ITensor(rng.random(2) + 0.1, (tn.linkind(e),)) for e in g.all_edges()}
    info = B.BeliefPropagationResult()
    cache = B.beliefpropagation(tn, init, stopping_criterion=dict(maxiter=200, tol=1e-12), message_update_algorithm=alg, info=info)
    assert info.delta < 1e-12 and info.iterations < 200
    cp = B.canonical_arrays(tn)
    p = oracle.make_problem(cp.ga, cp.tensors, "single")
    msgs = [init[cp.ga.named_edge(e)].data for e in range(cp.ga.ne)]
    out, it, delta = oracle.beliefpropagation(p, msgs, maxiter=200, tol=1e-12, schedule="jacobi")
    assert abs(it - info.iterations) <= 1  # (69 sweeps; the residual shrinks by 0.58 per sweep)
    assert np.isclose(B.bethe_free_energy(tn, cache), oracle.bethe_free_energy(p, out), rtol=1e-9)
    # exact value of the 4x4 torus for scale: the Bethe approximation is close but not equal on a loopy graph
    z = oracle.contract_all_sequential(p)
    assert abs(B.bethe_free_energy(tn, cache) - math.log(z)) / math.log(z) < 0.1  # 13.86 vs 14.56


def test_bench_ising_lattice_vertex_kernel_against_generic_and_sampled_oracle(oracle):
    """The bench workload's recipe on a 1024 x 640 torus (655 360 vertices: more than one pass of the kernel's grid-stride
    loop): VERTEX kernel vs the generic kernel on the same inputs, sampled edges vs the oracle, normalisation."""
    p = problems.synthetic_ising((1024, 640))
    ga = p.ga
    res = {}
    out = {}
    for kernel in (AUTO, GENERIC):
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p, kernel)
            assert all(b["kernel"] == (VERTEX if kernel == AUTO else GENERIC) for b in ctx.buckets())
            r = []
            for _ in range(2):
                r.append(ctx.sweep(1)[0])
            res[kernel], out[kernel] = r, ctx.get_messages_flat()
    assert np.abs(out[AUTO] - out[GENERIC]).max() < 1e-13
    assert np.allclose(res[AUTO], res[GENERIC], rtol=0, atol=1e-13)
    assert np.allclose(out[AUTO].reshape(-1, 2).sum(axis=1), 1.0, rtol=0, atol=1e-13)
    # one sweep from the initial messages, sampled edges recomputed by the oracle from the packed inputs
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(1)
        after = ctx.get_messages_flat().reshape(-1, 2)
    before = p.messages.reshape(-1, 2)
    rng = np.random.default_rng(3)
    worst = 0.0
    for e in rng.choice(ga.ne, size=64, replace=False):
        u = int(ga.src[e])
        lo, hi = int(ga.row_ptr[u]), int(ga.row_ptr[u + 1])
        T = p.tensors[16 * u:16 * (u + 1)].reshape((2,) * (hi - lo), order="F")
        ins = [None if f == e else before[ga.rev[f]] for f in range(lo, hi)]
        want = oracle.normalize_message(oracle.contract_single(T, int(ga.slot[e]), ins))
        worst = max(worst, np.abs(after[e] - want).max() / np.abs(want).max())
    assert worst < MSG_RTOL, worst
