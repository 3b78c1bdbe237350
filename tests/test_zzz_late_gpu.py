"""GPU tests written after the round's GPU budget had ended (first executed by the round-end run): kept in a file that
sorts after every test already seen green on a B200, so that `pytest -x` reaches those first."""
import numpy as np
import pytest

import itnn_b200 as B
from test_abi import _build_demo
from test_zz_golden import _apply_fixture, _check_apply_result, rel

pytestmark = pytest.mark.gpu


def test_gpu_reproduces_apply_fixture():
    f, ga, tensors, msgs, apply_state = _apply_fixture()
    with B.BPXContext(0) as ctx:
        ctx.set_graph(ga.src, ga.dst, ga.slot, ga.nv)
        ctx.set_dims(np.complex128, "norm", [2] * ga.nv, [3] * ga.ne)
        ctx.set_site_tensors(tensors)
        ctx.set_messages(msgs)
        svs = ctx.apply_two_site_gates([int(e) for e in f["edges"]], list(f["ops"]), max_rank=int(f["max_rank"]), normalize=True)
        shapes = [t.shape for t in tensors]
        new_tensors = [ctx.get_site_tensor(v).reshape(shapes[v], order="F") for v in range(ga.nv)]
        new_msgs = ctx.get_messages()
        _check_apply_result(f, ga, apply_state, new_tensors, new_msgs, svs)
        # the one-site gate of the fixture on the ORIGINAL state
        ctx.set_site_tensors(tensors)
        ctx.set_messages(msgs)
        v = int(f["one_site_vertex"])
        ctx.apply_one_site_gates([v], [f["one_site_op"]], normalize=True)
        assert rel(ctx.get_site_tensor(v), f["one_site_result"]) < 1e-10


def test_c_client_runs_on_the_gpu(pkg, tmp_path):
    """examples/bpx_demo.c: a plain-C99 client of include/bpx.h (BP sweeps, a gate, sweeps again, beliefs)."""
    r = _build_demo(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BP:" in r.stdout and "kept singular values" in r.stdout and "vertex scalar" in r.stdout


@pytest.mark.parametrize("schedule", ["synchronous", "sequential"])
def test_bp_ai_layer_gpu(schedule):
    """AI.solve / manual stepping / user-defined criteria over a device-resident iterate (tests/test_algorithmsinterface.py)."""
    from test_algorithmsinterface import check_bp_ai_layer

    check_bp_ai_layer(schedule)


def test_gate_layer_at_the_true_cfg5_shape():
    """Degree-4 vertices with chi = 16, d = 2 (1 MiB tensors, 4096 x 32 matrix views): the four inner vertices of a 4 x 4
    lattice, two disjoint gates in one layer, against the apply oracle through random probes on the external legs (the
    full pair product would be 0.5 GB)."""
    from helpers import randn
    from itnn_b200 import graphs, problems
    from oracle import apply_oracle as A
    from test_zz_gpu_apply import device_tensors, oracle_state

    rng = np.random.default_rng(5)
    chi = 16
    p = problems.synthetic_peps(graphs.named_grid((4, 4)), chi, 2, np.float64, init="positive")
    ga = p.ga
    vid = {v: i for i, v in enumerate(ga.vertices)}
    edges = [ga.edge_index[(vid[(2, 2)], vid[(3, 2)])], ga.edge_index[(vid[(3, 3)], vid[(2, 3)])]]
    ops = [np.eye(4).reshape(2, 2, 2, 2) + 0.3 * randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    with B.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(3, 0.0, True)
        msgs = ctx.get_messages()
        svs = ctx.apply_two_site_gates(edges, ops, max_rank=chi, normalize=True)
        new_tensors = device_tensors(ctx, p)
    state = oracle_state(p, p.tensors)
    env = {(ga.src[e], ga.dst[e]): msgs[e] for e in range(ga.ne)}

    def probe(t, slot, vecs):
        out = t
        for leg in reversed(range(t.ndim - 1)):
            if leg != slot:
                out = np.tensordot(out, vecs[leg], axes=([1 + leg], [0]))
        return out

    for e, op, sv in zip(edges, ops, svs):
        v1, v2, r = ga.src[e], ga.dst[e], ga.rev[e]
        assert new_tensors[v1].shape == (2, 16, 16, 16, 16) and new_tensors[v2].shape == (2, 16, 16, 16, 16)
        names = (("s", v1), ("s", v2))
        want_state, want_env = A.apply_operator((op, names, names), state, env, trunc=chi, normalize=True)
        assert np.allclose(sv, np.diag(want_env[(v1, v2)]).real, rtol=1e-8, atol=1e-12)
        w1 = A.permute(want_state[v1], state[v1][1])
        w2 = A.permute(want_state[v2], state[v2][1])
        p1 = [rng.standard_normal(chi) for _ in range(4)]
        p2 = [rng.standard_normal(chi) for _ in range(4)]
        got = probe(new_tensors[v1], ga.slot[e], p1) @ probe(new_tensors[v2], ga.slot[r], p2).T
        want = probe(w1, ga.slot[e], p1) @ probe(w2, ga.slot[r], p2).T
        assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()
