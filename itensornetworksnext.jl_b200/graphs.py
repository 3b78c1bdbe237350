"""Host-side graph containers and lattice generators for the BP hot path.

These mirror the *inputs* the reference obtains from NamedGraphs.jl (not vendored under
/root/reference): `named_grid`, `named_path_graph`, `named_cycle_graph`, `named_comb_tree`
(used at test/test_beliefpropagation.jl:157-225, test/test_apply_operator.jl:64,89), plus the
heavy-hex 127-site lattice that BASELINE.json config 3 names and the reference does not ship
(SURVEY.md F7).  Pure host bookkeeping: no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, Hashable, Iterable, List, Sequence, Tuple

Vertex = Hashable


@dataclass(frozen=True)
class NamedEdge:
    """Directed edge `src => dst` (NamedGraphs.NamedEdge analogue)."""

    src: Vertex
    dst: Vertex

    def reverse(self) -> "NamedEdge":
        return NamedEdge(self.dst, self.src)

    def __iter__(self):
        yield self.src
        yield self.dst

    def __repr__(self) -> str:
        return f"{self.src!r} => {self.dst!r}"


def to_edge(e) -> NamedEdge:
    if isinstance(e, NamedEdge):
        return e
    s, d = e
    return NamedEdge(s, d)


class NamedGraph:
    """Undirected simple graph with ordered vertices and ordered neighbour lists.

    Neighbour order is the order in which edges were added; `incident_edges(v)` returns the
    directed edges `v => w` in that order.  The leg order of site tensors built by
    `random_state` follows it (reference recipe: test/test_normnetwork.jl:22-31).
    """

    def __init__(self, vertices: Iterable[Vertex] = ()):  # noqa: D401
        self._adj: Dict[Vertex, List[Vertex]] = {}
        for v in vertices:
            self.add_vertex(v)

    # -- construction -------------------------------------------------------------------
    def add_vertex(self, v: Vertex) -> None:
        self._adj.setdefault(v, [])

    def add_edge(self, u: Vertex, v: Vertex) -> None:
        if u == v:
            raise ValueError("self loops are not supported")
        self.add_vertex(u)
        self.add_vertex(v)
        if v not in self._adj[u]:
            self._adj[u].append(v)
            self._adj[v].append(u)

    # -- queries ------------------------------------------------------------------------
    def vertices(self) -> List[Vertex]:
        return list(self._adj.keys())

    def neighbors(self, v: Vertex) -> List[Vertex]:
        return list(self._adj[v])

    def degree(self, v: Vertex) -> int:
        return len(self._adj[v])

    def has_edge(self, u: Vertex, v: Vertex) -> bool:
        return u in self._adj and v in self._adj[u]

    def edges(self) -> List[NamedEdge]:
        """Each undirected edge once, oriented from the earlier vertex to the later one."""
        pos = {v: i for i, v in enumerate(self._adj)}
        out = []
        for u in self._adj:
            for w in self._adj[u]:
                if pos[u] < pos[w]:
                    out.append(NamedEdge(u, w))
        return out

    def all_edges(self) -> List[NamedEdge]:
        """Both orientations of every edge (NamedGraphs `all_edges`), grouped by source."""
        return [NamedEdge(u, w) for u in self._adj for w in self._adj[u]]

    def incident_edges(self, v: Vertex) -> List[NamedEdge]:
        return [NamedEdge(v, w) for w in self._adj[v]]

    def in_incident_edges(self, v: Vertex) -> List[NamedEdge]:
        return [NamedEdge(w, v) for w in self._adj[v]]

    def nv(self) -> int:
        return len(self._adj)

    def ne(self) -> int:
        return sum(len(a) for a in self._adj.values()) // 2

    def is_tree(self) -> bool:
        return self.ne() == self.nv() - 1 and len(connected_components(self)) == 1


def connected_components(g: NamedGraph) -> List[List[Vertex]]:
    seen, comps = set(), []
    for r in g.vertices():
        if r in seen:
            continue
        comp, stack = [], [r]
        seen.add(r)
        while stack:
            v = stack.pop()
            comp.append(v)
            for w in g.neighbors(v):
                if w not in seen:
                    seen.add(w)
                    stack.append(w)
        comps.append(comp)
    return comps


# ---------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------
def named_grid(dims: Sequence[int] | int, periodic: bool = False) -> NamedGraph:
    """Hypercubic lattice with 1-based tuple vertices, first coordinate fastest.

    `named_grid((4, 4))`, `named_grid((16, 16, 16); periodic = true)` analogues.  Periodic
    wrap edges are only added along dimensions of length > 2 (no double edges).
    """
    if isinstance(dims, int):
        dims = (dims,)
    dims = tuple(int(d) for d in dims)
    verts = [tuple(reversed(c)) for c in itertools.product(*[range(1, d + 1) for d in reversed(dims)])]
    g = NamedGraph(verts)
    for v in verts:
        for ax, n in enumerate(dims):
            if v[ax] < n:
                w = v[:ax] + (v[ax] + 1,) + v[ax + 1:]
                g.add_edge(v, w)
            elif periodic and n > 2:
                w = v[:ax] + (1,) + v[ax + 1:]
                g.add_edge(v, w)
    return g


def named_path_graph(n: int) -> NamedGraph:
    g = NamedGraph(range(1, n + 1))
    for i in range(1, n):
        g.add_edge(i, i + 1)
    return g


def named_cycle_graph(n: int) -> NamedGraph:
    g = named_path_graph(n)
    if n > 2:
        g.add_edge(n, 1)
    return g


def named_comb_tree(dims: Tuple[int, int]) -> NamedGraph:
    """Comb tree: a backbone path of length dims[0] along y = 1, a tooth of length dims[1] on each."""
    nx, ny = dims
    verts = [(i, j) for j in range(1, ny + 1) for i in range(1, nx + 1)]
    g = NamedGraph(verts)
    for i in range(1, nx):
        g.add_edge((i, 1), (i + 1, 1))
    for i in range(1, nx + 1):
        for j in range(1, ny):
            g.add_edge((i, j), (i, j + 1))
    return g


def heavy_hex_127() -> NamedGraph:
    """127-site / 144-edge heavy-hex lattice (IBM Eagle topology), vertices 0..126.

    Seven horizontal chains (14, 15, 15, 15, 15, 15, 14 sites) joined by six groups of four
    bridge sites.  Authored here: the reference has no heavy-hex generator.
    """
    g = NamedGraph(range(127))
    rows = [
        list(range(0, 14)),
        list(range(18, 33)),
        list(range(37, 52)),
        list(range(56, 71)),
        list(range(75, 90)),
        list(range(94, 109)),
        list(range(113, 127)),
    ]
    for r in rows:
        for a, b in zip(r[:-1], r[1:]):
            g.add_edge(a, b)
    bridges = [
        ([14, 15, 16, 17], [0, 4, 8, 12], [18, 22, 26, 30]),
        ([33, 34, 35, 36], [20, 24, 28, 32], [39, 43, 47, 51]),
        ([52, 53, 54, 55], [37, 41, 45, 49], [56, 60, 64, 68]),
        ([71, 72, 73, 74], [58, 62, 66, 70], [77, 81, 85, 89]),
        ([90, 91, 92, 93], [75, 79, 83, 87], [94, 98, 102, 106]),
        ([109, 110, 111, 112], [96, 100, 104, 108], [114, 118, 122, 126]),
    ]
    for bs, ups, downs in bridges:
        for b, u, d in zip(bs, ups, downs):
            g.add_edge(u, b)
            g.add_edge(b, d)
    assert g.nv() == 127 and g.ne() == 144
    return g


# ---------------------------------------------------------------------------------------
# edge sequences
# ---------------------------------------------------------------------------------------
def forest_cover_edge_sequence(g: NamedGraph) -> List[NamedEdge]:
    """Default sequential BP schedule (reference: beliefpropagation.jl:14).

    Restates the published behaviour of NamedGraphs `forest_cover_edge_sequence` (dependency not
    vendored): split the edges into edge-disjoint spanning forests; for each tree
    emit the post-order DFS edges directed towards the root, then the same list reversed and
    flipped.  Every directed edge appears exactly once, and on a tree one sweep is exact.
    """
    remaining = {frozenset((e.src, e.dst)) for e in g.edges()}
    seq: List[NamedEdge] = []
    while remaining:
        # spanning forest of the remaining edge set
        used = set()
        seen = set()
        for root in g.vertices():
            if root in seen:
                continue
            if not any(frozenset((root, w)) in remaining for w in g.neighbors(root)):
                continue
            seen.add(root)
            # iterative DFS recording tree edges (parent -> child)
            order: List[Tuple[Vertex, Vertex]] = []
            stack = [(root, iter(g.neighbors(root)))]
            post: List[NamedEdge] = []
            while stack:
                v, it = stack[-1]
                advanced = False
                for w in it:
                    key = frozenset((v, w))
                    if key in remaining and key not in used and w not in seen:
                        seen.add(w)
                        used.add(key)
                        order.append((v, w))
                        stack.append((w, iter(g.neighbors(w))))
                        advanced = True
                        break
                if not advanced:
                    stack.pop()
                    if stack:
                        parent = stack[-1][0]
                        post.append(NamedEdge(v, parent))  # child -> parent, post-order
            seq.extend(post)
            seq.extend(e.reverse() for e in reversed(post))
        if not used:
            raise RuntimeError("forest cover failed to make progress")
        remaining -= used
    return seq


@dataclass
class GraphArrays:
    """Canonical integer view of a graph, the form the C ABI takes (include/bpx.h: bpx_set_graph).

    Vertices are numbered in `g.vertices()` order; directed edges are grouped by source (CSR) in
    neighbour order, so `slot[e]` (position of the link among the source's link legs) is the
    offset inside the source's row.
    """

    vertices: List[Vertex]
    vindex: Dict[Vertex, int]
    src: List[int]
    dst: List[int]
    rev: List[int]
    slot: List[int]
    row_ptr: List[int]
    edge_index: Dict[Tuple[int, int], int] = field(default_factory=dict)

    @property
    def nv(self) -> int:
        return len(self.vertices)

    @property
    def ne(self) -> int:
        return len(self.src)

    def edge_id(self, e) -> int:
        e = to_edge(e)
        return self.edge_index[(self.vindex[e.src], self.vindex[e.dst])]

    def named_edge(self, eid: int) -> NamedEdge:
        return NamedEdge(self.vertices[self.src[eid]], self.vertices[self.dst[eid]])


def graph_arrays(g: NamedGraph) -> GraphArrays:
    verts = g.vertices()
    vindex = {v: i for i, v in enumerate(verts)}
    src, dst, slot, row_ptr = [], [], [], [0]
    edge_index: Dict[Tuple[int, int], int] = {}
    for v in verts:
        for k, w in enumerate(g.neighbors(v)):
            edge_index[(vindex[v], vindex[w])] = len(src)
            src.append(vindex[v])
            dst.append(vindex[w])
            slot.append(k)
        row_ptr.append(len(src))
    rev = [edge_index[(d, s)] for s, d in zip(src, dst)]
    return GraphArrays(verts, vindex, src, dst, rev, slot, row_ptr, edge_index)


def grid_graph_arrays(dims: Sequence[int], periodic: bool = False) -> GraphArrays:
    """Hypercubic lattice straight to the integer arrays of the C ABI, vectorised (numpy): for lattices of millions of
    vertices, where building a `NamedGraph` of Python tuples first would take minutes.

    Vertex ids run with the first coordinate fastest (like `named_grid`); the link legs of a vertex are ordered
    (-axis0, +axis0, -axis1, +axis1, ...), existing neighbours only -- NOT the insertion order `named_grid` produces, so
    site tensors must be authored for this leg order.  Periodic wrap links only along axes longer than 2.  `vertices`
    is a `range`, and the name-based lookups (`vindex`, `edge_index`, `edge_id`) are not populated."""
    import numpy as np

    dims = tuple(int(d) for d in dims)
    nd = len(dims)
    nv = int(np.prod(dims))
    idx = np.arange(nv, dtype=np.int64)
    coords = np.unravel_index(idx, dims, order="F")
    nbr = np.full((nv, 2 * nd), -1, dtype=np.int64)
    stride = 1
    for ax, n in enumerate(dims):
        c = coords[ax]
        wrap = periodic and n > 2
        nbr[:, 2 * ax] = np.where(c > 0, idx - stride, idx + (n - 1) * stride if wrap else -1)
        nbr[:, 2 * ax + 1] = np.where(c < n - 1, idx + stride, idx - (n - 1) * stride if wrap else -1)
        stride *= n
    valid = nbr >= 0
    deg = valid.sum(axis=1)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    pos = np.cumsum(valid, axis=1) - 1            # position of direction k among the existing legs of a vertex
    src = np.repeat(idx, deg)
    dst = nbr[valid]
    slot = pos[valid]
    direction = np.broadcast_to(np.arange(2 * nd), nbr.shape)[valid]
    rev = row_ptr[dst] + pos[dst, direction ^ 1]  # the same link seen from the other end: opposite direction there
    return GraphArrays(range(nv), {}, src, dst, rev, slot.astype(np.int64), row_ptr, {})

