"""Vertex partitioning of a BP problem over ranks (one process per GPU) and the host-side plumbing that
connects the ranks' libbpx contexts.

The data path needs no collective library: after `connect`, every sweep pushes the messages on cut edges
straight into the peer ranks' message buffers over NVLink peer memory and exchanges the residual through
peer mailboxes (csrc/bpx_halo.cuh).  torch.distributed is used ONLY to hand the CUDA IPC handles around.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np


@dataclass
class HaloPlan:
    """What rank `rank` sends and receives per sweep (pure bookkeeping, mirrors bpx_set_partition)."""

    rank: int
    owned_vertices: List[int]
    owned_edges: List[int]              # directed edges whose source is owned: the updates this rank performs
    send: Dict[int, List[int]]          # peer -> owned edges whose head lives on that peer
    recv: Dict[int, List[int]]          # peer -> that peer's edges pointing into an owned vertex


def plan(src: Sequence[int], dst: Sequence[int], owner: Sequence[int], rank: int) -> HaloPlan:
    owned_v = [v for v, o in enumerate(owner) if o == rank]
    owned_e, send, recv = [], {}, {}
    for e, (u, v) in enumerate(zip(src, dst)):
        if owner[u] == rank:
            owned_e.append(e)
            if owner[v] != rank:
                send.setdefault(owner[v], []).append(e)
        elif owner[v] == rank:
            recv.setdefault(owner[u], []).append(e)
    return HaloPlan(rank, owned_v, owned_e, send, recv)


def strip_owner(vertices: Sequence, nranks: int, axis: int = -1) -> List[int]:
    """Contiguous slabs along one lattice axis (vertices are 1-based coordinate tuples)."""
    coords = [v[axis] for v in vertices]
    lo, hi = min(coords), max(coords)
    n = hi - lo + 1
    return [min(nranks - 1, (c - lo) * nranks // n) for c in coords]


def block_owner(nv: int, nranks: int) -> List[int]:
    """Balanced contiguous ranges of the vertex order (general graphs)."""
    return [min(nranks - 1, v * nranks // nv) for v in range(nv)]


def connect(ctx, owner: Sequence[int], rank: int, world: int, group=None) -> None:
    """bpx_set_partition + exchange of the IPC handles + bpx_halo_connect for every peer."""
    import torch.distributed as dist

    own = np.ascontiguousarray(owner, dtype=np.int32)
    ctx._check(ctx.lib.bpx_set_partition(ctx.h, int(rank), int(world), own.ctypes.data_as(C.c_void_p)))
    if world == 1:
        return
    buf = (C.c_ubyte * 192)()
    ctx._check(ctx.lib.bpx_halo_export(ctx.h, buf))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf), group=group)
    for peer in range(world):
        if peer == rank:
            continue
        hb = (C.c_ubyte * 192).from_buffer_copy(handles[peer])
        ctx._check(ctx.lib.bpx_halo_connect(ctx.h, peer, hb))
    dist.barrier(group=group)  # nobody sweeps before every rank has opened its peers' buffers
