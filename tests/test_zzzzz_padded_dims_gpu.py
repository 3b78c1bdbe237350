"""Link dimensions 9..15 on degree-4 Float64 vertices (csrc/bpx_pad.cuh): the library zero-pads those links to 16 inside a child
context so that the chi = 16 tensor-pipe kernels serve them; every result at the C ABI must equal the unpadded problem's --
sweeps and residuals against the oracle, beliefs, host iterates, tensors in and out, gate layers against the unpadded
context (`BPX_NO_PAD=1`, generic update kernel)."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import peps_tensors, randn, rel_err
from itnn_b200 import _lib, graphs, problems
from test_gpu_parity import MSG_RTOL, make_ctx
from test_zz_gpu_apply import bond_invariant, matching, oracle_state

pytestmark = pytest.mark.gpu


def positive_messages(rng, link_dim):
    out = []
    for c in link_dim:
        m = np.eye(c) + 0.1 * np.abs(rng.standard_normal((c, c)))
        out.append(np.asfortranarray(m / m.sum()))
    return out


def lattice(dims, chi_of_edge, rng):
    ga = graphs.graph_arrays(graphs.named_grid(dims))
    link = [0] * ga.ne
    for e in range(ga.ne):
        if e < ga.rev[e]:
            link[e] = link[ga.rev[e]] = chi_of_edge(e)
    tensors = peps_tensors(ga, None, 2, np.float64, rng, link_dim=[link[f] for f in range(ga.ne)])
    return ga, link, tensors


CASES = {
    "chi12": lambda e: 12,
    "chi10": lambda e: 10,
    "mixed_9_to_15": lambda e: 9 + e % 7,
    "chi12_and_16": lambda e: 12 if e % 3 else 16,
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_padded_sweeps_beliefs_and_host_iterates_match_the_oracle(oracle, case):
    rng = np.random.default_rng(5)
    ga, link, tensors = lattice((4, 5), CASES[case], rng)
    msgs = positive_messages(rng, link)
    p = oracle.make_problem(ga, tensors, "norm")
    with make_ctx(ga, np.float64, "norm", [2] * ga.nv, link, tensors, msgs) as ctx:
        kinds = {b["degree"]: (b["kernel"], b["chi"]) for b in ctx.buckets()}
        assert kinds[4] == (_lib.BPX_KERNEL_SLICED, 16)          # the interior runs on the chi = 16 tensor-pipe kernel
        want = list(msgs)
        for k in range(3):
            prev, want = want, oracle.sweep_jacobi(p, want, True)
            res, done = ctx.sweep(1, 0.0, True)
            got = ctx.get_messages()
            assert [g.shape for g in got] == [(c, c) for c in link]
            assert rel_err(got, want) < MSG_RTOL, f"sweep {k}"
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
        assert np.allclose(ctx.vertex_scalars(), oracle.vertex_scalars(p, want), rtol=1e-10)
        assert np.allclose(ctx.edge_scalars(), oracle.edge_scalars(p, want), rtol=1e-10)
        f = ctx.bethe_free_energy()
        assert abs(f - oracle.bethe_free_energy(p, want)) <= 1e-10 * abs(f)
        assert abs(ctx.iterate_diff(prev) - oracle.iterate_diff(want, prev)) < 1e-11
        # tensors come back in the caller's dims; a single message too
        for v in (0, ga.nv // 2, ga.nv - 1):
            assert np.array_equal(ctx.get_site_tensor(v), np.asarray(tensors[v]).ravel(order="F"))
        # host iterate in, one sweep, host iterate out
        flat_in = np.concatenate([np.asarray(m).ravel(order="F") for m in want])
        flat_out = np.empty_like(flat_in)
        res = ctx.sweep_host(flat_in, flat_out)
        nxt = oracle.sweep_jacobi(p, want, True)
        assert rel_err(ctx.unpack_messages(flat_out), nxt) < MSG_RTOL
        assert abs(res - oracle.iterate_diff(nxt, want)) < 1e-11
        # BP to convergence: same number of sweeps as the oracle
        ctx.set_messages(msgs)
        res, done = ctx.sweep(200, 1e-10)
        _, it, _ = oracle.beliefpropagation(p, msgs, maxiter=200, tol=1e-10)
        assert done == it and res < 1e-10


def test_no_pad_switch_and_forced_policy(monkeypatch, oracle):
    rng = np.random.default_rng(6)
    ga, link, tensors = lattice((4, 4), CASES["chi12"], rng)
    msgs = positive_messages(rng, link)
    monkeypatch.setenv("BPX_NO_PAD", "1")
    with make_ctx(ga, np.float64, "norm", [2] * ga.nv, link, tensors, msgs) as ctx:
        kinds = {b["degree"]: (b["kernel"], b["chi"]) for b in ctx.buckets()}
        assert kinds[4] == (_lib.BPX_KERNEL_GENERIC, 12)
        res0, _ = ctx.sweep(1, 0.0, True)
        ref = ctx.get_messages()
    monkeypatch.delenv("BPX_NO_PAD")
    with make_ctx(ga, np.float64, "norm", [2] * ga.nv, link, tensors, msgs) as ctx:
        res1, _ = ctx.sweep(1, 0.0, True)
        assert rel_err(ctx.get_messages(), ref) < MSG_RTOL and abs(res0 - res1) < 1e-12
    with make_ctx(ga, np.float64, "norm", [2] * ga.nv, link, tensors, msgs, kernel=_lib.BPX_KERNEL_GENERIC) as ctx:
        assert {b["degree"]: b["chi"] for b in ctx.buckets()}[4] == 12  # a forced kernel policy keeps the caller's dims


@pytest.mark.parametrize("max_rank,normalize", [(0, True), (7, False)])
def test_gate_layer_on_a_padded_context(monkeypatch, max_rank, normalize):
    """Gates see zero-padded tensors and messages: the Gram-path kernel takes them (exactly zero message rows are padding,
    not rank deficiency) and the result equals the unpadded context's."""
    rng = np.random.default_rng(9)
    ga, link, tensors = lattice((4, 4), CASES["chi10"], rng)
    msgs = positive_messages(rng, link)
    p = problems.SyntheticProblem("", ga, np.dtype(np.float64), 10, 2, [2] * ga.nv, link, tensors, msgs)
    edges = matching(ga, rng)
    ops = [np.eye(4).reshape(2, 2, 2, 2) + 0.2 * randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    out = []
    for no_pad in ("1", None):
        if no_pad:
            monkeypatch.setenv("BPX_NO_PAD", no_pad)
        else:
            monkeypatch.delenv("BPX_NO_PAD")
        with make_ctx(ga, np.float64, "norm", [2] * ga.nv, link, tensors, msgs) as ctx:
            ctx.sweep(4, 0.0, True)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=normalize)
            stats = ctx.apply_stats()
            shapes = [(2,) + tuple(link[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1])) for v in range(ga.nv)]
            ts = [ctx.get_site_tensor(v).reshape(shapes[v], order="F") for v in range(ga.nv)]
            res, _ = ctx.sweep(1, 0.0, True)
            out.append((svs, ts, res, stats))
    (sv_a, t_a, r_a, st_a), (sv_b, t_b, r_b, st_b) = out
    assert st_a == (len(edges), 0) and st_b == (len(edges), 0)
    s_a, s_b = oracle_state(p, t_a), oracle_state(p, t_b)
    for e, a, b in zip(edges, sv_a, sv_b):
        assert a.shape == b.shape == (link[e],)
        assert np.allclose(a, b, rtol=1e-9, atol=1e-13)
        x, y = bond_invariant(s_a, ga.src[e], ga.dst[e]), bond_invariant(s_b, ga.src[e], ga.dst[e])
        assert np.abs(x - y).max() <= 1e-9 * np.abs(x).max()
    assert abs(r_a - r_b) < 1e-9


def test_padded_synthetic_data_is_zero_outside_the_callers_dims():
    q = problems.synthetic_peps(graphs.named_grid((4, 4)), 12, 2, np.float64, host_data=False)
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, q)
        kinds = {b["degree"]: (b["kernel"], b["chi"]) for b in ctx.buckets()}
        assert kinds[4] == (_lib.BPX_KERNEL_SLICED, 16)
        res, done = ctx.sweep(200, 1e-10)
        assert res < 1e-10 and done < 200
        m = ctx.get_messages()
        assert all(x.shape == (12, 12) for x in m) and all(abs(x.sum() - 1.0) < 1e-12 for x in m)
        assert np.isfinite(ctx.bethe_free_energy())


def test_multi_device_context_pads_too(oracle):
    """A context over two devices: the padded child is a multi-device context over the same device list (the per-device
    contexts themselves never pad).  Needs two GPUs (two sliced kernels cannot share one device)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(3)
    ga, link, tensors = lattice((6, 6), CASES["chi12"], rng)
    msgs = positive_messages(rng, link)
    p = oracle.make_problem(ga, tensors, "norm")
    with B.BPXContext(devices=[0, 1]) as ctx:
        ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        ctx.set_dims(np.float64, "norm", [2] * ga.nv, link)
        assert ctx.lib.bpx_num_devices(ctx.h) == 2
        kinds = {b["degree"]: (b["kernel"], b["chi"]) for b in ctx.buckets()}
        assert kinds[4] == (_lib.BPX_KERNEL_SLICED, 16)
        ctx.set_site_tensors(tensors)
        ctx.set_messages(msgs)
        want = list(msgs)
        for k in range(2):
            prev, want = want, oracle.sweep_jacobi(p, want, True)
            res, done = ctx.sweep(1, 0.0, True)
            assert rel_err(ctx.get_messages(), want) < MSG_RTOL
            assert abs(res - oracle.iterate_diff(want, prev)) < 1e-11
        assert np.allclose(ctx.vertex_scalars(), oracle.vertex_scalars(p, want), rtol=1e-10)
        f = ctx.bethe_free_energy()
        assert abs(f - oracle.bethe_free_energy(p, want)) <= 1e-10 * abs(f)
        assert np.array_equal(ctx.get_site_tensor(ga.nv // 2), np.asarray(tensors[ga.nv // 2]).ravel(order="F"))
        # re-declaring the dims (here: unpadded shape 8) turns the handle back into a plain two-device context
        link8 = [8] * ga.ne
        ctx.set_dims(np.float64, "norm", [2] * ga.nv, link8)
        assert ctx.lib.bpx_num_devices(ctx.h) == 2
        assert {b["degree"]: b["chi"] for b in ctx.buckets()}[4] == 8
