"""`bethe_free_energy` reduced on the device (bpx_bethe_free_energy; messagecache.jl:185-201) against the
oracle: plain real result, the complex-promotion rule (a negative term in a real network; complex element types), the
-Inf rule, and the partial sums of multi-device contexts."""
import math

import numpy as np
import pytest

import itnn_b200 as B
from helpers import single_layer_tensors
from itnn_b200 import graphs, problems
from test_gpu_parity import make_ctx

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _close(got, want, rtol=1e-10):
    if isinstance(want, complex) or np.iscomplexobj(want):
        # the imaginary part is a sum of phases: compare modulo 2 pi
        assert isinstance(got, complex)
        assert abs(got.real - want.real) <= rtol * max(1.0, abs(want.real))
        d = (got.imag - want.imag) % (2 * math.pi)
        assert min(d, 2 * math.pi - d) < 1e-8
    else:
        assert isinstance(got, float)
        assert abs(got - want) <= rtol * max(1.0, abs(want))


@pytest.mark.parametrize("name,dims,periodic", [("cfg1", (4, 4), False), ("cfg2", (5, 4), False), ("cfg4", (3, 3, 3), True),
                                                ("cfg5", (4, 4), False), ("cfg2c", (4, 3), False), ("cfg3", None, False)])
def test_norm_network_configs(oracle, name, dims, periodic):
    p = problems.make_config(name, graph=graphs.named_grid(dims, periodic=periodic)) if dims else problems.make_config(name)
    op = oracle.make_problem(p.ga, p.tensors, "norm")
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(3)
        msgs = ctx.get_messages()
        want = oracle.bethe_free_energy(op, msgs)
        got = ctx.bethe_free_energy()
        _close(got, want if np.iscomplexobj(want) else float(want))
        # the message sets are untouched by the belief pass
        assert all(np.array_equal(a, b) for a, b in zip(ctx.get_messages(), msgs))
        # second call: cached work space, same bits
        again = ctx.bethe_free_energy()
        assert again == got


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_signed_single_layer_network_promotes_like_the_reference(oracle, dtype):
    """Random signed factors: some vertex / edge scalars are negative, so the reference promotes the terms to complex and
    the result carries i*pi per negative term (messagecache.jl:189-194)."""
    rng = np.random.default_rng(11)
    ga = graphs.graph_arrays(graphs.named_grid((4, 4), periodic=True))
    tensors = single_layer_tensors(ga, 2, dtype, rng)
    msgs = [rng.standard_normal(2).astype(dtype) for _ in range(ga.ne)]
    p = oracle.make_problem(ga, tensors, "single")
    vs = np.asarray(oracle.vertex_scalars(p, msgs))
    es = np.asarray(oracle.edge_scalars(p, msgs))
    if np.dtype(dtype).kind != "c":
        assert (vs < 0).any() or (es < 0).any()  # the case under test
    want = oracle.bethe_free_energy(p, msgs)
    with make_ctx(ga, dtype, "single", None, [2] * ga.ne, tensors, msgs) as ctx:
        got = ctx.bethe_free_energy()
        _close(got, complex(want))
        parts = ctx.bethe_free_energy_parts()
        assert parts[4] == float((vs.real < 0).any()) and parts[5] == float((es.real < 0).any()) and parts[6] == 0.0


def test_zero_edge_scalar_gives_minus_infinity(oracle):
    rng = np.random.default_rng(3)
    ga = graphs.graph_arrays(graphs.named_path_graph(4))
    tensors = single_layer_tensors(ga, 2, np.float64, rng)
    msgs = [rng.random(2) + 0.1 for _ in range(ga.ne)]
    e = 0
    r = int(ga.rev[e])
    msgs[e] = np.array([1.0, 0.0])
    msgs[r] = np.array([0.0, 1.0])  # <m_e, m_rev(e)> = 0
    p = oracle.make_problem(ga, tensors, "single")
    assert oracle.bethe_free_energy(p, msgs) == -np.inf
    with make_ctx(ga, np.float64, "single", None, [2] * ga.ne, tensors, msgs) as ctx:
        got = ctx.bethe_free_energy()
        assert got == -math.inf


@pytest.mark.parametrize("devices", [[0, 0], [0, 1]])
def test_multi_device_partial_sums(oracle, devices):
    if max(devices) >= _ngpu():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    p = problems.make_config("cfg2", graph=graphs.named_grid((6, 5)))
    with B.BPXContext(0) as one:
        problems.upload(one, p)
        one.sweep(2)
        want = one.bethe_free_energy()
    with B.BPXContext(devices=devices) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(2)
        got = ctx.bethe_free_energy()
    assert isinstance(got, float) and abs(got - want) <= 1e-12 * abs(want)
