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


def test_a_cache_keeps_its_own_values_when_its_session_moves_on(oracle):
    """Two host caches share one device session (beliefpropagation.py `_session_for`): after the session has swept for the
    second cache, beliefs of the FIRST cache must still be those of the first cache's messages (the session is
    tagged with a version and re-uploads the cache that asks)."""
    import itnn_b200 as B
    from itnn_b200 import graphs

    g = graphs.named_grid((3, 3))
    tn, _, _ = B.random_state(np.float64, g, d=2, chi=3, rng=np.random.default_rng(5))
    nn = B.normnetwork(tn)
    env0 = B.message_environment(B.ones_message, nn)
    env1 = B.beliefpropagation(nn, env0, stopping_criterion=dict(maxiter=1), message_update_algorithm=B.B200MessageUpdate())
    f1 = B.bethe_free_energy(nn, env1)
    vs1 = np.array(B.vertex_scalars(nn, env1))
    # advance the SAME session through another cache (single-edge updates re-use env1's session)
    env2 = B.MessageCache(dict(env1.items()))
    env2._session, env2._session_version = env1._session, env1._session_version
    e = next(iter(dict(env2.items())))
    for _ in range(3):
        B.message_update(env2, nn, e)
    assert env2._session is env1._session and env1._session_version != env1._session.version
    assert np.allclose(np.array(B.vertex_scalars(nn, env1)), vs1, rtol=1e-13)
    assert abs(B.bethe_free_energy(nn, env1) - f1) <= 1e-12 * abs(f1)
