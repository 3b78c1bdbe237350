"""Host-side mirror of `ITensorNetworksNext.ITensorNetworkGenerators` -- the INPUTS of the BP path.

  diagonaltensor / delta ... src/ITensorNetworkGenerators/delta_network.jl:6-36
  delta_network ............ src/ITensorNetworkGenerators/delta_network.jl:38-53
  sqrt_ising_bond .......... src/ITensorNetworkGenerators/ising_network.jl:7-17
  ising_network ............ src/ITensorNetworkGenerators/ising_network.jl:19-51

These build single-layer `ITensorNetwork`s (one-off host bookkeeping: a handful of 2x2 matrices and
delta tensors); they perform no BP arithmetic.  The networks they return are consumed unchanged by
`beliefpropagation` (single-layer mode of the device path, include/bpx.h BPX_MODE_SINGLE).

`f(edge)` must return the link `Index` of the edge for either orientation, as in the reference's
tests (`l(e) = get(() -> ldict[reverse(e)], ldict, e)`, test/test_itensornetworkgenerators.jl:17-18).
"""
from __future__ import annotations

from typing import Callable, Hashable, Iterable, Sequence

import numpy as np

from .graphs import NamedEdge, NamedGraph
from .tensornetwork import Index, ITensor, ITensorNetwork, tensornetwork, uniquename


def diagonaltensor(diag: Sequence, inds: Sequence[Index]) -> ITensor:
    """Zero tensor over `inds` with `diag` on its main diagonal (delta_network.jl:22-34)."""
    diag = np.asarray(diag)
    shape = tuple(i.dim for i in inds)
    a = np.zeros(shape, dtype=diag.dtype, order="F")
    n = min(shape) if shape else 1
    if len(diag) < n:
        raise ValueError(f"diagonal of length {len(diag)} is shorter than the tensor's diagonal ({n})")
    for k in range(n):
        a[(k,) * len(shape)] = diag[k]
    return ITensor(a, inds)


def delta(dtype, inds: Sequence[Index]) -> ITensor:
    """delta(elt, is) = diagonaltensor(ones(elt, min dim), is) (delta_network.jl:36)."""
    n = min((i.dim for i in inds), default=1)
    return diagonaltensor(np.ones(n, dtype=np.dtype(dtype)), inds)


def delta_network(f: Callable[[NamedEdge], Index], g: NamedGraph, dtype=np.float64) -> ITensorNetwork:
    """A delta (copy) tensor on every vertex over the link indices `f.(incident_edges(g, v))`
    (delta_network.jl:45-53)."""
    return tensornetwork(lambda v: delta(dtype, [f(e) for e in g.incident_edges(v)]), g.vertices())


def sqrt_ising_bond(beta: float, J: float = 1.0, h: float = 0.0, *, deg1: int, deg2: int) -> np.ndarray:
    """Square root (through the eigendecomposition, `v * sqrt(D) * inv(v)`) of the 2x2 Boltzmann bond
    matrix with the field split evenly over the bonds of each end point (ising_network.jl:7-17)."""
    h1, h2 = h / deg1, h / deg2
    m = np.array(
        [
            [np.exp(beta * (J + h1 + h2)), np.exp(beta * (-J + h1 - h2))],
            [np.exp(beta * (-J - h1 + h2)), np.exp(beta * (J - h1 - h2))],
        ]
    )
    d, v = np.linalg.eig(m)
    if np.iscomplexobj(d) or np.any(d < 0):
        # Julia: sqrt of a negative real eigenvalue throws DomainError (antiferromagnetic J needs a complex beta)
        raise ValueError("sqrt_ising_bond: bond matrix has a negative or complex eigenvalue (DomainError in the reference)")
    return v @ np.diag(np.sqrt(d)) @ np.linalg.inv(v)


def ising_network(f: Callable[[NamedEdge], Index], beta: float, g: NamedGraph, J: float = 1.0, h: float = 0.0,
                  sz_vertices: Iterable[Hashable] = ()) -> ITensorNetwork:
    """Ising partition-function network on `g` at inverse temperature `beta` (ising_network.jl:27-51):
    a delta tensor per vertex (diag(1, -1) on `sz_vertices`, which inserts sigma^z there), and on every
    bond the square root of the Boltzmann matrix absorbed into BOTH end points:
        T_v[..., l, ...] <- sum_{l~} m[l~, l] T_v[..., l~, ...]."""
    dtype = np.result_type(type(beta), np.float64)
    tilde = {}
    for e in g.edges():
        l = f(e)
        tilde[frozenset((e.src, e.dst))] = Index(l.dim, uniquename())

    def fp(e):
        return tilde[frozenset((e.src, e.dst))]

    tensors = {v: delta(dtype, [fp(e) for e in g.incident_edges(v)]) for v in g.vertices()}
    for v in sz_vertices:
        tensors[v] = diagonaltensor(np.array([1, -1], dtype=dtype), tensors[v].inds)
    for e in g.edges():
        l, lt = f(e), fp(e)
        if l.dim != 2:
            raise ValueError("ising_network: link indices must have dimension 2")
        m = sqrt_ising_bond(beta, J, h, deg1=g.degree(e.src), deg2=g.degree(e.dst)).astype(dtype)
        for v in (e.src, e.dst):
            t = tensors[v]
            ax = t.dimnames().index(lt.name)
            data = np.moveaxis(np.tensordot(t.data, m, axes=([ax], [0])), -1, ax)  # contract l~, new leg l in place
            inds = list(t.inds)
            inds[ax] = l
            tensors[v] = ITensor(np.asfortranarray(data), inds)
    return ITensorNetwork(tensors)
