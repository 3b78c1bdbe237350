"""Synthetic inputs for BASELINE.json's configurations, in the canonical layout of include/bpx.h.

Recipe: site tensors i.i.d. standard normal from the library's counter-based RNG
(`bpx_fill_randn`, seed 123 -- the seed of test/test_beliefpropagation.jl:155 -- stream = vertex id),
rescaled by (d * prod chi)^(-1/2); initial messages M = I + 0.1 |randn| sum-normalised (stream =
nv + edge id), or all ones (test/test_apply_operator.jl:72).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .device import fill_randn
from .graphs import GraphArrays, NamedGraph, graph_arrays, heavy_hex_127, named_grid


@dataclass
class SyntheticProblem:
    name: str
    ga: GraphArrays
    dtype: np.dtype
    chi: int
    d: int
    phys_dim: List[int]
    link_dim: List[int]
    tensors: List[np.ndarray]   # (d, chi, ..., chi), Fortran order  (single mode: (chi, ..., chi); or ONE packed 1-D array)
    messages: List[np.ndarray]  # (chi, chi) [bra, ket]             (single mode: (chi,);          or ONE packed 1-D array)
    mode: str = "norm"          # "norm": PEPS norm network (double layer) | "single": plain ITensorNetwork factors

    @property
    def n_updates(self) -> int:
        return self.ga.ne

    def flops_per_sweep(self) -> float:
        """Algorithmic flops: 2 z d chi^(z+1) per update, x4 for complex.  Single-layer networks:
        absorb the z-1 vector messages one leg at a time, 2 (chi^z + chi^(z-1) + ... + chi^2) per update."""
        c = 4.0 if self.dtype.kind == "c" else 1.0
        deg = np.diff(np.asarray(self.ga.row_ptr))
        if self.mode == "single":
            return float(sum(n * z * single_layer_update_flops(z, self.chi) for z, n in zip(*np.unique(deg, return_counts=True)))) * c
        tot = 0.0
        for v in range(self.ga.nv):
            z = self.ga.row_ptr[v + 1] - self.ga.row_ptr[v]
            tot += z * (2.0 * z * self.d * float(self.chi) ** (z + 1)) * c
        return tot

    def bytes_per_sweep(self) -> float:
        """Algorithmic bytes: every site tensor once + 3 x every message (in, old, out)."""
        w = self.dtype.itemsize
        if self.mode == "single":
            deg = np.diff(np.asarray(self.ga.row_ptr))
            return float((float(self.chi) ** deg).sum()) * w + 3.0 * self.ga.ne * self.chi * w
        site = sum(self.d * self.chi ** int(self.ga.row_ptr[v + 1] - self.ga.row_ptr[v]) for v in range(self.ga.nv)) * w
        return site + 3.0 * self.ga.ne * self.chi * self.chi * w


def single_layer_update_flops(z: int, chi: int) -> float:
    """One single-layer update in absorption order: the factor shrinks by chi with every absorbed message."""
    return 2.0 * sum(float(chi) ** k for k in range(2, z + 1)) if z >= 2 else 0.0


def synthetic_ising(dims, beta: float = 0.3, periodic: bool = True, seed: int = 123, name: str = "ising") -> SyntheticProblem:
    """Single-layer Ising partition-function network on a hypercubic lattice (the `ising_network` generator's tensors,
    src/ITensorNetworkGenerators/ising_network.jl:27-51, J = 1, h = 0), built straight in the packed layout for lattices
    of millions of vertices: T_v[l_0..l_{z-1}] = sum_s prod_k W[s, l_k] with W = sqrt of the 2x2 Boltzmann bond matrix.
    Initial messages: uniform (0.1, 1.1) random vectors from the shared RNG (the spin-ice test draws `rand`,
    test/test_beliefpropagation.jl:214-216), sum-normalised."""
    from .generators import sqrt_ising_bond
    from .graphs import grid_graph_arrays

    ga = grid_graph_arrays(dims, periodic)
    w = sqrt_ising_bond(beta, deg1=2, deg2=2)  # h = 0: the matrix does not depend on the degrees
    deg = np.diff(ga.row_ptr)
    by_deg = {}
    for z in np.unique(deg):
        t = np.zeros((2,) * int(z))
        for s in range(2):
            v = np.ones(())
            for _ in range(int(z)):
                v = np.multiply.outer(v, w[s])
            t = t + v
        by_deg[int(z)] = t.ravel(order="F")
    site_off = np.concatenate([[0], np.cumsum(2 ** deg.astype(np.int64))])
    sites = np.empty(int(site_off[-1]))
    for z, t in by_deg.items():
        vs = np.nonzero(deg == z)[0]
        sites[(site_off[vs][:, None] + np.arange(t.size)[None, :]).ravel()] = np.tile(t, len(vs))
    r = fill_randn(seed, ga.nv, np.float64, 2 * ga.ne)
    m = (0.1 + (np.abs(r) % 1.0)).reshape(ga.ne, 2)
    msgs = (m / m.sum(axis=1, keepdims=True)).ravel()
    return SyntheticProblem(name, ga, np.dtype(np.float64), 2, 1, [1] * 0, np.full(ga.ne, 2, dtype=np.int32), sites, msgs, "single")


def unpacked(p: SyntheticProblem):
    """(tensors, messages) of a problem as per-vertex / per-edge arrays (small problems: oracle comparisons)."""
    if not isinstance(p.tensors, np.ndarray):
        return p.tensors, p.messages
    deg = np.diff(np.asarray(p.ga.row_ptr))
    off = np.concatenate([[0], np.cumsum(p.chi ** deg.astype(np.int64))])
    tensors = [p.tensors[off[v]:off[v + 1]].reshape((p.chi,) * int(deg[v]), order="F") for v in range(p.ga.nv)]
    msgs = [p.messages[p.chi * e:p.chi * (e + 1)].copy() for e in range(p.ga.ne)]
    return tensors, msgs


def synthetic_peps(g: NamedGraph, chi: int, d: int = 2, dtype=np.float64, seed: int = 123, init: str = "positive",
                   name: str = "", host_data: bool = True) -> SyntheticProblem:
    """`host_data = False` builds the structure only (inputs are then generated on the device by
    `BPXContext.fill_synthetic`, same recipe) -- for workloads the host cannot stage (cfg5: 63 GiB)."""
    dtype = np.dtype(dtype)
    ga = graph_arrays(g)
    if not host_data:
        return SyntheticProblem(name, ga, dtype, chi, d, [d] * ga.nv, [chi] * ga.ne, None, None)
    tensors = []
    for v in range(ga.nv):
        z = ga.row_ptr[v + 1] - ga.row_ptr[v]
        shape = (d,) + (chi,) * z
        n = int(np.prod(shape))
        t = fill_randn(seed, v, dtype, n) * (1.0 / np.sqrt(n))
        tensors.append(t.reshape(shape, order="F"))
    msgs = []
    for e in range(ga.ne):
        if init == "ones":
            m = np.ones((chi, chi), dtype=dtype)
        elif init == "positive":
            r = fill_randn(seed, ga.nv + e, np.float64, chi * chi).reshape((chi, chi), order="F")
            m = (np.eye(chi) + 0.1 * np.abs(r)).astype(dtype)
            m = m / m.sum()
        elif init == "randn":
            m = fill_randn(seed, ga.nv + e, dtype, chi * chi).reshape((chi, chi), order="F")
        else:
            raise ValueError(init)
        msgs.append(np.asfortranarray(m))
    return SyntheticProblem(name, ga, dtype, chi, d, [d] * ga.nv, [chi] * ga.ne, tensors, msgs)


CONFIGS = {
    # BASELINE.json `configs`, in order
    "cfg1": dict(graph=lambda: named_grid((4, 4)), chi=2, d=2, dtype=np.float64),
    "cfg2": dict(graph=lambda: named_grid((32, 32)), chi=8, d=2, dtype=np.float64),
    "cfg3": dict(graph=heavy_hex_127, chi=16, d=2, dtype=np.complex128),
    "cfg4": dict(graph=lambda: named_grid((16, 16, 16), periodic=True), chi=4, d=2, dtype=np.float64),
    "cfg5": dict(graph=lambda: named_grid((256, 256)), chi=16, d=2, dtype=np.float64),
    # not a BASELINE config: the ComplexF64 twin of cfg2 (complex PEPS on the square lattice)
    "cfg2c": dict(graph=lambda: named_grid((32, 32)), chi=8, d=2, dtype=np.complex128),
}


def make_config(name: str, seed: int = 123, init: str = "positive", graph: Optional[NamedGraph] = None,
                host_data: bool = True) -> SyntheticProblem:
    if name == "ising":  # not a BASELINE config: the HBM-bound single-layer bucket, 1024x1024 periodic
        return synthetic_ising((1024, 1024), seed=seed)
    c = CONFIGS[name]
    g = c["graph"]() if graph is None else graph
    return synthetic_peps(g, c["chi"], c["d"], c["dtype"], seed, init, name, host_data)


def upload(ctx, p: SyntheticProblem, kernel: Optional[int] = None):
    """Describe `p` to a BPXContext and make site tensors and messages resident."""
    ctx.set_graph(p.ga.src, p.ga.dst, p.ga.slot, p.ga.nv)
    if kernel is not None:
        ctx.set_kernel_policy(kernel)
    ctx.set_dims(p.dtype, p.mode, p.phys_dim if p.mode == "norm" else None, p.link_dim)
    if p.tensors is None:
        ctx.fill_synthetic(123)
    else:
        ctx.set_site_tensors(p.tensors)
        ctx.set_messages(p.messages)
    return ctx
