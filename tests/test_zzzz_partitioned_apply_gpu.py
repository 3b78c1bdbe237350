"""Gate layers on a vertex-PARTITIONED run (2 GPUs): every rank applies the gates inside its own block between sweeps;
a gate across the cut is refused.  Written without a GPU (needs 2 devices; skipped otherwise) -- sorts last."""
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_partition import ROOT, _free_port, _problem, _setup

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _local_gates(ga, owner, rank, rng):
    """A matching inside the block of `rank`, plus one directed edge across the cut."""
    used, edges, cut = set(), [], None
    for e in rng.permutation(ga.ne):
        s, d = ga.src[e], ga.dst[e]
        if owner[s] == rank and owner[d] == rank and s not in used and d not in used:
            used |= {s, d}
            edges.append(int(e))
        elif owner[s] == rank and owner[d] != rank and cut is None:
            cut = int(e)
    return edges, cut


def _worker(rank, world, port, q):
    pkg, o = _setup(rank, world, port, "gloo")
    from helpers import randn
    from itnn_b200 import partition, problems

    torch.cuda.set_device(rank)
    g, p = _problem(pkg, dims=(6, 8), chi=4)
    ga = p.ga
    owner = partition.strip_owner(ga.vertices, world, axis=1)
    pl = partition.plan(ga.src, ga.dst, owner, rank)
    rng = np.random.default_rng(100 + rank)
    edges, cut = _local_gates(ga, owner, rank, rng)
    ops = [np.eye(4).reshape(2, 2, 2, 2) + 0.3 * randn(rng, np.float64, (2, 2, 2, 2)) for _ in edges]
    ctx = pkg.BPXContext(rank)
    problems.upload(ctx, p)
    partition.connect(ctx, owner, rank, world)
    ctx.sweep(3, 0.0)
    before = ctx.get_messages()
    refused = False
    try:
        ctx.apply_two_site_gates([cut], [ops[0]], max_rank=4)
    except pkg.BPXError as ex:
        refused = ex.status == -4 and "cut edge" in str(ex)
    svs = ctx.apply_two_site_gates(edges, ops, max_rank=4, normalize=True)
    shapes = {v: (2,) + tuple(p.link_dim[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1])) for v in pl.owned_vertices}
    tensors = {v: ctx.get_site_tensor(v).reshape(shapes[v], order="F") for v in pl.owned_vertices}
    after = ctx.get_messages()
    res, _ = ctx.sweep(1, 0.0)
    swept = ctx.get_messages()
    q.put((rank, refused, edges, ops, svs, tensors, {e: before[e] for e in range(ga.ne)}, {e: after[e] for e in pl.owned_edges},
           {e: swept[e] for e in pl.owned_edges}, res))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_two_rank_local_gate_layers():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, o = entry.import_package(), entry.import_oracle()
    from oracle import apply_oracle as A
    from test_zz_gpu_apply import bond_invariant, oracle_state

    g, p = _problem(pkg, dims=(6, 8), chi=4)
    ga = p.ga
    # messages after 3 sweeps: both ranks agree with the single-process oracle on what they read
    op = o.make_problem(ga, p.tensors, "norm")
    want = list(p.messages)
    for _ in range(3):
        want = o.sweep_jacobi(op, want)
    state = oracle_state(p, p.tensors)
    env = {(ga.src[e], ga.dst[e]): want[e] for e in range(ga.ne)}
    new_tensors = [None] * ga.nv
    new_msgs = list(want)
    for rank, refused, edges, ops, svs, tensors, before, after, swept, res in results:
        assert refused, "a gate across the cut must be refused with BPX_ERR_UNSUPPORTED"
        for v, t in tensors.items():
            new_tensors[v] = t
        for e, m in after.items():
            new_msgs[e] = m
    got_state = oracle_state(p, new_tensors)
    for rank, refused, edges, ops, svs, tensors, before, after, swept, res in results:
        for e, gate, sv in zip(edges, ops, svs):
            v1, v2 = ga.src[e], ga.dst[e]
            names = (("s", v1), ("s", v2))
            want_state, want_env = A.apply_operator((gate, names, names), state, env, trunc=4, normalize=True)
            assert np.allclose(sv, np.diag(want_env[(v1, v2)]).real, rtol=1e-8, atol=1e-12)
            x, y = bond_invariant(got_state, v1, v2), bond_invariant(want_state, v1, v2)
            assert np.abs(x - y).max() <= 1e-8 * np.abs(y).max()
    # the sweep after the gates, from the devices' own new tensors and messages
    op2 = o.make_problem(ga, new_tensors, "norm")
    want_swept = o.sweep_jacobi(op2, new_msgs)
    for rank, refused, edges, ops, svs, tensors, before, after, swept, res in results:
        for e, m in swept.items():
            assert np.abs(m - want_swept[e]).max() <= 1e-10 * np.abs(want_swept[e]).max()
        assert abs(res - o.iterate_diff(want_swept, new_msgs)) < 1e-11
