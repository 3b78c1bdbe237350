"""Vertex-partitioned BP: host-side plan + a world_size-2 run.

CPU (gloo): the partition plan of itnn_b200.partition drives a 2-process run in which every rank updates
only its own edges (with the oracle standing in for the kernels), ships the cut-edge messages and
max-reduces the residual -- the result must equal the single-process sweep.
GPU (marked `gpu`, needs 2 devices): the same through libbpx with the NVLink peer-memory halo push."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port, backend):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    import __graft_entry__ as entry

    return entry.import_package(), entry.import_oracle()


def _problem(pkg, dims=(6, 4), chi=3, dtype=np.float64):
    from itnn_b200 import graphs, problems

    g = graphs.named_grid(dims)
    return g, problems.synthetic_peps(g, chi, 2, dtype, seed=7)


def _worker_cpu(rank, world, port, nsweeps, q):
    pkg, o = _setup(rank, world, port, "gloo")
    from itnn_b200 import partition

    g, p = _problem(pkg)
    owner = partition.strip_owner(p.ga.vertices, world, axis=0)
    pl = partition.plan(p.ga.src, p.ga.dst, owner, rank)
    op = o.make_problem(p.ga, p.tensors, "norm")
    msgs = [m.copy() for m in p.messages]
    res_hist = []
    for _ in range(nsweeps):
        new = o.sweep_jacobi(op, msgs, edges=pl.owned_edges)  # only this rank's updates
        local = max(o.edge_residual(msgs[e], new[e]) for e in pl.owned_edges)
        # halo exchange: cut-edge messages to the rank that owns their head
        for peer in range(world):
            if peer == rank:
                continue
            out = [torch.from_numpy(np.ascontiguousarray(new[e])) for e in pl.send.get(peer, [])]
            inc = [torch.empty(p.chi, p.chi, dtype=torch.float64) for _ in pl.recv.get(peer, [])]
            reqs = [dist.isend(t, peer) for t in out] + [dist.irecv(t, peer) for t in inc]
            for r in reqs:
                r.wait()
            for e, t in zip(pl.recv.get(peer, []), inc):
                new[e] = t.numpy().copy()
        t = torch.tensor([local], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res_hist.append(float(t.item()))
        msgs = new
    # every rank must hold the correct value of every message it reads: own edges + incoming cut edges
    need = set(pl.owned_edges) | {e for es in pl.recv.values() for e in es}
    q.put((rank, {e: msgs[e] for e in need}, res_hist, pl))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_matches_single_process_gloo():
    world, nsweeps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_cpu, args=(r, world, port, nsweeps, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, o = entry.import_package(), entry.import_oracle()
    g, p = _problem(pkg)
    op = o.make_problem(p.ga, p.tensors, "norm")
    want = list(p.messages)
    hist = []
    for _ in range(nsweeps):
        prev, want = want, o.sweep_jacobi(op, want)
        hist.append(o.iterate_diff(want, prev))
    covered = set()
    for rank, msgs, res_hist, pl in results:
        assert np.allclose(res_hist, hist, rtol=0, atol=1e-14)
        for e, m in msgs.items():
            assert np.allclose(m, want[e], rtol=1e-13, atol=0)
        covered |= set(pl.owned_edges)
        # plan sanity: what one rank sends is what the other expects
        for peer, es in pl.send.items():
            other = next(r for r in results if r[0] == peer)[3]
            assert other.recv[rank] == es
    assert covered == set(range(p.ga.ne))


def test_plan_counts_strips():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    entry.import_package()
    from itnn_b200 import graphs, partition

    g = graphs.named_grid((32, 64))
    ga = graphs.graph_arrays(g)
    owner = partition.strip_owner(ga.vertices, 2, axis=1)
    pl0, pl1 = (partition.plan(ga.src, ga.dst, owner, r) for r in range(2))
    assert len(pl0.owned_vertices) == len(pl1.owned_vertices) == 1024
    assert len(pl0.send[1]) == len(pl1.send[0]) == 32  # one cut edge per lattice column in each direction
    assert len(pl0.owned_edges) + len(pl1.owned_edges) == ga.ne
    assert partition.block_owner(10, 3) == [0, 0, 0, 0, 1, 1, 1, 2, 2, 2]


# ---------------------------------------------------------------------------------------------------
def _worker_gpu(rank, world, port, nsweeps, q, chi=8):
    pkg, o = _setup(rank, world, port, "gloo")
    from itnn_b200 import partition, problems

    torch.cuda.set_device(rank)
    g, p = _problem(pkg, dims=(8, 12), chi=chi)
    owner = partition.strip_owner(p.ga.vertices, world, axis=1)
    ctx = pkg.BPXContext(rank)
    problems.upload(ctx, p)
    partition.connect(ctx, owner, rank, world)
    hist = []
    for _ in range(nsweeps):
        res, done = ctx.sweep(1, 0.0)
        hist.append(res)
    got = ctx.get_messages()
    res_tol, done_tol = ctx.sweep(50, 1e-3)  # convergence test on the GLOBAL residual
    pl = partition.plan(p.ga.src, p.ga.dst, owner, rank)
    # site tensors live on the owning rank only; vertex scalars are reported for owned vertices
    vs = ctx.vertex_scalars()
    final = ctx.get_messages()
    foreign = next(v for v in range(p.ga.nv) if owner[v] != rank)
    try:
        ctx.get_site_tensor(foreign)
        resident_foreign = True
    except KeyError:
        resident_foreign = False
    assert not resident_foreign
    assert np.allclose(ctx.get_site_tensor(pl.owned_vertices[0]), p.tensors[pl.owned_vertices[0]].ravel(order="F"))
    op = o.make_problem(p.ga, p.tensors, "norm")
    for v in pl.owned_vertices[:6]:
        # incoming messages of an owned vertex are valid on this rank (own edges + received cut edges)
        assert np.isclose(vs[v], o.vertex_scalar(op, final, v), rtol=1e-10)
    assert all(vs[v] == 0 for v in range(p.ga.nv) if owner[v] != rank)
    need = set(pl.owned_edges) | {e for es in pl.recv.values() for e in es}
    q.put((rank, {e: got[e] for e in need}, hist, (res_tol, done_tol), ctx.buckets()))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("chi", [8, 12])  # 12: every rank's context zero-pads its links to 16 (csrc/bpx_pad.cuh)
def test_two_rank_partition_gpu_peer_halo(chi):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, nsweeps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gpu, args=(r, world, port, nsweeps, q, chi)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, o = entry.import_package(), entry.import_oracle()
    g, p = _problem(pkg, dims=(8, 12), chi=chi)
    op = o.make_problem(p.ga, p.tensors, "norm")
    want = list(p.messages)
    hist = []
    for _ in range(nsweeps):
        prev, want = want, o.sweep_jacobi(op, want)
        hist.append(o.iterate_diff(want, prev))
    _, it, delta = o.beliefpropagation(op, want, maxiter=50, tol=1e-3)
    for rank, msgs, h, (res_tol, done_tol), buckets in results:
        assert np.allclose(h, hist, rtol=0, atol=1e-12)
        for e, m in msgs.items():
            assert np.abs(m - want[e]).max() / np.abs(want[e]).max() < 1e-10
        assert done_tol == it and abs(res_tol - delta) < 1e-11


# ---------------------------------------------------------------------------------------------------
def _worker_gpu_sweep_host(rank, world, port, nsteps, q):
    """bpx_sweep_host on a partitioned context: every rank uploads / receives only the messages it owns."""
    pkg, o = _setup(rank, world, port, "gloo")
    from itnn_b200 import partition, problems

    torch.cuda.set_device(rank)
    g, p = _problem(pkg, dims=(8, 12), chi=8)
    owner = partition.strip_owner(p.ga.vertices, world, axis=1)
    pl = partition.plan(p.ga.src, p.ga.dst, owner, rank)
    out = {}
    for mode in ("pinned", "pageable"):
        ctx = pkg.BPXContext(rank)
        problems.upload(ctx, p)
        partition.connect(ctx, owner, rank, world)
        flat = ctx.pack_messages(p.messages)
        if mode == "pinned":  # streamed: kernel gated on the chunked upload, results stored straight into host memory
            ta, tb = torch.empty(flat.size, dtype=torch.float64).pin_memory(), torch.empty(flat.size, dtype=torch.float64).pin_memory()
            a, b = ta.numpy(), tb.numpy()
        else:
            a, b = np.empty_like(flat), np.empty_like(flat)
        a[:] = flat
        b[:] = np.nan
        hist = []
        for _ in range(nsteps):
            hist.append(ctx.sweep_host(a, b))
            a, b = b, a
        msgs = ctx.unpack_messages(a)
        out[mode] = ({e: msgs[e] for e in pl.owned_edges}, hist)
        dist.barrier()
        ctx.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_sweep_host_moves_owned_messages_only():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, nsteps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gpu_sweep_host, args=(r, world, port, nsteps, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, o = entry.import_package(), entry.import_oracle()
    g, p = _problem(pkg, dims=(8, 12), chi=8)
    op = o.make_problem(p.ga, p.tensors, "norm")
    want = list(p.messages)
    hist = []
    for _ in range(nsteps):
        prev, want = want, o.sweep_jacobi(op, want)
        hist.append(o.iterate_diff(want, prev))
    seen = set()
    for rank, out in results:
        for mode in ("pinned", "pageable"):
            msgs, h = out[mode]
            assert np.allclose(h, hist, rtol=0, atol=1e-12), (rank, mode)  # the GLOBAL residual on every rank
            for e, m in msgs.items():
                assert np.abs(m - want[e]).max() / np.abs(want[e]).max() < 1e-10, (rank, mode, e)
        seen |= set(out["pinned"][0])
    assert seen == set(range(p.ga.ne))  # the ranks' owned slices tile the iterate


# ---------------------------------------------------------------------------------------------------
def _worker_gpu_multi_launch(rank, world, port, nsweeps, q):
    """chi = 16 lattice: degree 4 (sliced kernel), degree 3 (tuned 16-wide kernel) and degree 2 (16-wide slice kernel) are
    three launches per sweep; the exchange is fused into them (first launch gates, last launch posts)."""
    pkg, o = _setup(rank, world, port, "gloo")
    from itnn_b200 import partition, problems

    torch.cuda.set_device(rank)
    g, p = _problem(pkg, dims=(4, 6), chi=16)
    owner = partition.strip_owner(p.ga.vertices, world, axis=1)
    pl = partition.plan(p.ga.src, p.ga.dst, owner, rank)
    ctx = pkg.BPXContext(rank)
    problems.upload(ctx, p)
    partition.connect(ctx, owner, rank, world)
    hist = []
    for _ in range(nsweeps):
        res, done = ctx.sweep(1, 0.0)
        hist.append(res)
    got = ctx.get_messages()
    launches = ctx.counters()["launches"]
    # and the streamed host step over the same three launches (pinned buffers)
    flat = ctx.pack_messages(p.messages)
    ta, tb = torch.empty(flat.size, dtype=torch.float64).pin_memory(), torch.empty(flat.size, dtype=torch.float64).pin_memory()
    a, b = ta.numpy(), tb.numpy()
    a[:] = flat
    ctx.set_messages(flat)
    hist_host = []
    for _ in range(nsweeps):
        hist_host.append(ctx.sweep_host(a, b))
        a, b = b, a
    host_msgs = ctx.unpack_messages(a)
    need = set(pl.owned_edges) | {e for es in pl.recv.values() for e in es}
    q.put((rank, {e: got[e] for e in need}, hist, {e: host_msgs[e] for e in pl.owned_edges}, hist_host, ctx.buckets(), launches))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_multi_launch_sweep_fuses_the_exchange():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, nsweeps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gpu_multi_launch, args=(r, world, port, nsweeps, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, o = entry.import_package(), entry.import_oracle()
    g, p = _problem(pkg, dims=(4, 6), chi=16)
    op = o.make_problem(p.ga, p.tensors, "norm")
    want = list(p.messages)
    hist = []
    for _ in range(nsweeps):
        prev, want = want, o.sweep_jacobi(op, want)
        hist.append(o.iterate_diff(want, prev))
    for rank, msgs, h, host_msgs, h_host, buckets, launches in results:
        assert sorted(b["kernel"] for b in buckets if b["edges"]) == [2, 2, 3], buckets  # on-chip x2 + sliced, no generic launch
        assert np.allclose(h, hist, rtol=0, atol=1e-12)
        assert np.allclose(h_host, hist, rtol=0, atol=1e-12)
        for e, m in msgs.items():
            assert np.abs(m - want[e]).max() / np.abs(want[e]).max() < 1e-10
        for e, m in host_msgs.items():
            assert np.abs(m - want[e]).max() / np.abs(want[e]).max() < 1e-10
