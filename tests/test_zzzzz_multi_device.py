"""Single-process multi-GPU contexts (bpx_create_multi): ONE context over a device list must give the
single-device result bit for bit -- same per-vertex arithmetic, only the owner of each vertex changes.

`devices = [0, 0]` puts both children on one GPU (small lattices whose launches fit side by side), so the protocol is
covered on the 1-GPU driver box; `[0, 1]` needs two devices.  CPU part: the entry points exist and fail loudly without a
device (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest

from itnn_b200 import _lib, graphs, problems
import itnn_b200 as B


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def test_create_multi_fails_loudly_without_a_device():
    if _ngpu() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(B.BPXError) as ei:
        B.BPXContext(devices=[0, 1])
    assert "no CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)


def test_create_multi_rejects_bad_arguments():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.bpx_create_multi(None, 2, C.byref(h)) == -1
    devs = np.zeros(1, dtype=np.int32)
    assert lib.bpx_create_multi(devs.ctypes.data_as(C.c_void_p), 0, C.byref(h)) == -1
    assert lib.bpx_num_devices(None) == -1


def _run(ctx, p, nsweeps=3):
    problems.upload(ctx, p)
    hist = []
    for _ in range(nsweeps):
        res, done = ctx.sweep(1)
        hist.append(res)
    msgs = ctx.get_messages_flat()
    vs = ctx.vertex_scalars()
    es = ctx.edge_scalars()
    return msgs, np.array(hist), vs, es


CASES = [("cfg2", (6, 6), False), ("cfg1", (4, 4), False), ("cfg4", (4, 4, 4), True), ("cfg2c", (5, 4), False), ("cfg5", (4, 4), False)]


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0, 0], [0, 1], [0, 0, 0]])
@pytest.mark.parametrize("name,dims,periodic", CASES)
def test_multi_context_equals_single_device_bit_for_bit(devices, name, dims, periodic):
    if max(devices) >= _ngpu():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    if name == "cfg5" and len(set(devices)) < len(devices):
        pytest.skip("the sliced kernel fills the device: children need their own GPUs")
    g = graphs.named_grid(dims, periodic=periodic)
    p = problems.make_config(name, graph=g)
    with B.BPXContext(0) as one:
        want = _run(one, p)
    with B.BPXContext(devices=devices) as ctx:
        assert ctx.lib.bpx_num_devices(ctx.h) == len(devices)
        got = _run(ctx, p)
        owner = ctx.get_owner()
        assert set(owner.tolist()) == set(range(len(devices)))          # every device owns vertices
        assert ctx.lib.bpx_num_cut_edges(ctx.h) > 0
        # one message, one site tensor through the owner lookup
        e = p.ga.ne // 2
        m = np.empty(int(ctx.msg_off[e + 1] - ctx.msg_off[e]), dtype=ctx.dtype)
        ctx._check(ctx.lib.bpx_get_message(ctx.h, e, m.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(m, got[0][ctx.msg_off[e]:ctx.msg_off[e + 1]])
        v = p.ga.nv - 1
        assert np.array_equal(ctx.get_site_tensor(v), np.asarray(p.tensors[v]).ravel(order="F"))
    if name == "cfg5":
        # the group-cooperative sliced kernel sums a vertex's partial tiles in the order of ITS group: the vertex-to-group
        # map changes with the partition, so results agree to rounding, not bit for bit
        assert np.abs(got[0] - want[0]).max() <= 1e-14 * np.abs(want[0]).max()
        assert np.allclose(got[1], want[1], rtol=0, atol=1e-13)
        assert np.allclose(got[2], want[2], rtol=1e-12) and np.allclose(got[3], want[3], rtol=1e-12)
    else:
        assert np.array_equal(got[0], want[0])          # messages after three sweeps
        assert np.array_equal(got[1], want[1])          # global residual of every sweep
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0, 0], [0, 1]])
def test_multi_context_converges_and_streams_like_single_device(devices):
    if max(devices) >= _ngpu():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    p = problems.make_config("cfg2", graph=graphs.named_grid((6, 6)))
    with B.BPXContext(0) as one:
        problems.upload(one, p)
        res1, done1 = one.sweep(200, 1e-10)
        want = one.get_messages_flat()
        hist1 = one.residual_history()
    with B.BPXContext(devices=devices) as ctx:
        problems.upload(ctx, p)
        res2, done2 = ctx.sweep(200, 1e-10)
        assert (done2, res2) == (done1, res1)          # StopWhenConverged fires after the same sweep
        assert np.array_equal(ctx.get_messages_flat(), want)
        assert np.array_equal(ctx.residual_history()[:done1], hist1[:done1])
        # explicit owner: two stripes the other way round
        own = ctx.get_owner()
        ctx.set_owner((len(devices) - 1 - own).astype(np.int32))
        ctx.set_site_tensors(p.tensors)
        ctx.set_messages(p.messages)
        res3, done3 = ctx.sweep(200, 1e-10)
        assert (done3, res3) == (done1, res1)
        assert np.array_equal(ctx.get_messages_flat(), want)
        # host iterates: one call per sweep, every device moves its own messages
        a = ctx.pack_messages(p.messages)
        b = np.empty_like(a)
        with B.BPXContext(0) as ref:
            problems.upload(ref, p)
            a1, b1 = a.copy(), np.empty_like(a)
            for _ in range(3):
                r_ref = ref.sweep_host(a1, b1)
                r_got = ctx.sweep_host(a, b)
                assert r_ref == r_got
                assert np.array_equal(b, b1)
                a, b, a1, b1 = b, a, b1, a1


@pytest.mark.gpu
def test_multi_context_refuses_what_it_cannot_do():
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    p = problems.make_config("cfg2", graph=graphs.named_grid((6, 6)))
    with B.BPXContext(devices=[0, 0]) as ctx:
        problems.upload(ctx, p)
        with pytest.raises(B.BPXError):
            ctx.set_stream(12345)
        own = np.zeros(p.ga.nv, dtype=np.int32)
        rc = ctx.lib.bpx_set_partition(ctx.h, 0, 2, own.ctypes.data_as(C.c_void_p))
        assert rc == -1 and b"multi-device" in ctx.lib.bpx_last_error(ctx.h)
        with pytest.raises(B.BPXError):
            ctx.set_owner(np.full(p.ga.nv, 7, dtype=np.int32))
