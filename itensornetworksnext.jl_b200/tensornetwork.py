"""Host-side mirror of the reference's network containers, as far as the BP path reads them.

  Index / ITensor ........ named-dimension arrays (ITensorBase.jl, not vendored) -- storage + names only
  ITensorNetwork ......... src/tensornetwork.jl:21-31, 115-173 (edges inferred from shared index names)
  NormNetwork ............ src/normnetwork.jl:13-93 (ket network + ket->bra link-name map)
  KetView / BraView ...... src/normnetworkview.jl:6-41
  random_state ........... test/test_normnetwork.jl:22-31 (the random-PEPS recipe)

These classes hold data and names; they perform NO contraction.  `canonical_arrays` lowers a network to
the flat layout of include/bpx.h, which is what the device path consumes.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Callable, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from .graphs import GraphArrays, NamedEdge, NamedGraph, graph_arrays, to_edge

_uid = itertools.count(1)


def uniquename() -> str:
    return f"i{next(_uid)}"


@dataclass(frozen=True)
class Index:
    dim: int
    name: Hashable = None

    def __post_init__(self):
        if self.name is None:
            object.__setattr__(self, "name", uniquename())

    def __len__(self):
        return self.dim


class ITensor:
    """A dense array with one `Index` per axis.  No arithmetic beyond elementwise helpers."""

    def __init__(self, data, inds: Sequence[Index]):
        data = np.asarray(data)
        inds = tuple(inds)
        if data.shape != tuple(i.dim for i in inds):
            raise ValueError(f"shape {data.shape} does not match indices {[i.dim for i in inds]}")
        if len({i.name for i in inds}) != len(inds):
            raise ValueError("repeated index name")
        self.data = data
        self.inds = inds

    @property
    def dtype(self):
        return self.data.dtype

    def dimnames(self):
        return tuple(i.name for i in self.inds)

    def array(self, *names) -> np.ndarray:
        """The data with axes permuted to the given name order."""
        have = self.dimnames()
        if set(names) != set(have) or len(names) != len(have):
            raise ValueError(f"names {names} do not match {have}")
        return np.transpose(self.data, [have.index(n) for n in names])

    def replacedimnames(self, f: Callable) -> "ITensor":
        return ITensor(self.data, [Index(i.dim, f(i.name)) for i in self.inds])

    def conj(self) -> "ITensor":
        return ITensor(np.conj(self.data), self.inds)

    def copy(self) -> "ITensor":
        return ITensor(self.data.copy(), self.inds)

    def __repr__(self):
        return f"ITensor({self.data.shape}, names={self.dimnames()}, {self.dtype})"


def randn_itensor(rng: np.random.Generator, dtype, inds: Sequence[Index]) -> ITensor:
    shape = tuple(i.dim for i in inds)
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        data = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)
    else:
        data = rng.standard_normal(shape)
    return ITensor(data.astype(dtype), inds)


class ITensorNetwork:
    """vertex -> ITensor; two vertices are joined when their tensors share an index name
    (src/tensornetwork.jl:138-173).  A name held by one tensor only is a site index."""

    def __init__(self, tensors: Dict[Hashable, ITensor]):
        self.tensors: Dict[Hashable, ITensor] = dict(tensors)
        self.dimname_vertices: Dict[Hashable, List[Hashable]] = {}
        for v, t in self.tensors.items():
            for n in t.dimnames():
                self.dimname_vertices.setdefault(n, []).append(v)
        self.graph = NamedGraph(self.tensors.keys())
        self._linkname: Dict[Tuple[Hashable, Hashable], Hashable] = {}
        for v, t in self.tensors.items():  # neighbour order = order of the link legs in the tensor
            for n in t.dimnames():
                vs = self.dimname_vertices[n]
                if len(vs) > 2:
                    raise ValueError(f"index {n!r} is shared by more than two tensors")
                if len(vs) == 2:
                    w = vs[0] if vs[1] == v else vs[1]
                    if (v, w) in self._linkname:
                        raise ValueError(f"vertices {v!r} and {w!r} share more than one index")
                    self._linkname[(v, w)] = n
                    self.graph._adj[v].append(w)

    def __getitem__(self, v) -> ITensor:
        return self.tensors[v]

    def vertices(self):
        return self.graph.vertices()

    def edges(self):
        return self.graph.edges()

    def linkname(self, edge) -> Hashable:
        e = to_edge(edge)
        try:
            return self._linkname[(e.src, e.dst)]
        except KeyError:
            raise KeyError(f"no link between {e.src!r} and {e.dst!r}") from None

    def linkind(self, edge) -> Index:
        e = to_edge(edge)
        n = self.linkname(e)
        return next(i for i in self.tensors[e.src].inds if i.name == n)

    def sitenames(self, v) -> List[Hashable]:
        return [n for n in self.tensors[v].dimnames() if len(self.dimname_vertices[n]) == 1]

    def has_dimname(self, name) -> bool:
        return name in self.dimname_vertices

    @property
    def dtype(self):
        return np.result_type(*[t.dtype for t in self.tensors.values()])


def tensornetwork(f: Callable, vertices) -> ITensorNetwork:
    return ITensorNetwork({v: f(v) for v in vertices})


class NormNetwork:
    """Double-layer <tn|tn> (src/normnetwork.jl:13-28): the ket network plus `braname[ketlink] = bralink`.
    The factor at v is ket_v ⊗ conj(bra_v) with the site legs shared (normnetwork.jl:49-54); it is never
    materialised here -- the kernels consume the ket tensor and conjugate on the fly."""

    def __init__(self, ket: ITensorNetwork, braname: Optional[Dict] = None):
        self.ket = ket
        links = [n for n, vs in ket.dimname_vertices.items() if len(vs) == 2]
        if braname is None:
            braname = {n: uniquename() for n in links}
        self._braname = {n: braname[n] for n in links}

    def vertices(self):
        return self.ket.vertices()

    def edges(self):
        return self.ket.edges()

    @property
    def graph(self) -> NamedGraph:
        return self.ket.graph

    def braname(self, name):
        if not self.ket.has_dimname(name):
            raise KeyError(f"index name {name} not found underlying tensor network.")
        return self._braname.get(name, name)  # site names map to themselves (normnetwork.jl:66-73)

    def kettensor(self, v) -> ITensor:
        return self.ket[v]

    def conj_bratensor(self, v) -> ITensor:
        return self.ket[v].replacedimnames(self.braname)

    def bratensor(self, v) -> ITensor:
        return self.conj_bratensor(v).conj()

    @property
    def dtype(self):
        return self.ket.dtype


def normnetwork(tn: ITensorNetwork, braname=None) -> NormNetwork:
    return NormNetwork(tn, braname)


class KetView:
    def __init__(self, nn: NormNetwork):
        self.parent = nn

    def __getitem__(self, v):
        return self.parent.kettensor(v)

    def linkname(self, edge):
        return self.parent.ket.linkname(edge)


class BraView:
    def __init__(self, nn: NormNetwork):
        self.parent = nn

    def __getitem__(self, v):
        return self.parent.bratensor(v)

    def linkname(self, edge):
        return self.parent.braname(self.parent.ket.linkname(edge))


def random_state(dtype, g: NamedGraph, d: int = 2, chi: int = 2, rng=None):
    """Random PEPS on `g` (test/test_normnetwork.jl:22-31): site index first, then one link index per
    incident edge in `incident_edges` order; i.i.d. standard normal entries."""
    rng = np.random.default_rng(123) if rng is None else rng
    l = {}
    for e in g.edges():
        l[frozenset((e.src, e.dst))] = Index(chi)
    s = {v: Index(d) for v in g.vertices()}

    def make(v):
        inds = [s[v]] + [l[frozenset((e.src, e.dst))] for e in g.incident_edges(v)]
        return randn_itensor(rng, dtype, inds)

    return tensornetwork(make, g.vertices()), l, s


# ---------------------------------------------------------------------------------------------------
# lowering to the canonical layout of include/bpx.h
# ---------------------------------------------------------------------------------------------------
@dataclass
class CanonicalProblem:
    ga: GraphArrays
    mode: str                 # "norm" | "single"
    dtype: np.dtype
    phys_dim: List[int]
    link_dim: List[int]       # per directed edge
    tensors: List[np.ndarray]  # norm: (d, chi_0..chi_{z-1}); single: (chi_0..chi_{z-1})
    ket_names: List[Hashable]  # per directed edge: link name in the ket layer
    bra_names: List[Hashable]  # per directed edge: link name in the bra layer (== ket name in single mode)


def canonical_arrays(factors) -> CanonicalProblem:
    """Walk `vertices(nn)`, `kettensor(nn, v)` and permute each tensor to [sites..., links in neighbour
    order] (SURVEY.md §8 b2 iv).  Several site legs on one vertex are fused into one."""
    if isinstance(factors, NormNetwork):
        mode, net = "norm", factors.ket
    elif isinstance(factors, ITensorNetwork):
        mode, net = "single", factors
    else:
        raise TypeError(f"unsupported factor container {type(factors).__name__}")
    ga = graph_arrays(net.graph)
    dtype = np.dtype(np.complex128 if np.dtype(net.dtype).kind == "c" else np.float64)
    tensors, phys = [], []
    for v in ga.vertices:
        t = net[v]
        sites = net.sitenames(v)
        links = [net.linkname(NamedEdge(v, w)) for w in net.graph.neighbors(v)]
        if mode == "single" and sites:
            raise ValueError(f"vertex {v!r} has uncontracted site legs; wrap the state in a NormNetwork")
        arr = t.array(*sites, *links).astype(dtype)
        if mode == "norm":
            d = int(np.prod([arr.shape[i] for i in range(len(sites))], dtype=np.int64)) if sites else 1
            arr = arr.reshape((d,) + arr.shape[len(sites):])
            phys.append(d)
        else:
            phys.append(1)
        tensors.append(np.asfortranarray(arr))
    ket_names, bra_names, link_dim = [], [], []
    for e in range(ga.ne):
        ne_ = ga.named_edge(e)
        kn = net.linkname(ne_)
        ket_names.append(kn)
        bra_names.append(factors.braname(kn) if mode == "norm" else kn)
        link_dim.append(net.linkind(ne_).dim)
    return CanonicalProblem(ga, mode, dtype, phys, link_dim, tensors, ket_names, bra_names)
