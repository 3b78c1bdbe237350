"""Shared builders for the tests (host-side only)."""
from __future__ import annotations

import numpy as np

from itnn_b200 import graphs


def randn(rng, dtype, shape):
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def single_layer_tensors(ga, chi, dtype, rng):
    """Random single-layer factors T_v[l_0..l_{z-1}] (test/test_beliefpropagation.jl:163-166)."""
    return [randn(rng, dtype, (chi,) * (ga.row_ptr[v + 1] - ga.row_ptr[v])) for v in range(ga.nv)]


def peps_tensors(ga, chi, d, dtype, rng, link_dim=None):
    """Random PEPS A_v[s, l_0..l_{z-1}] (test/test_normnetwork.jl:22-31); optional per-edge link dims."""
    out = []
    for v in range(ga.nv):
        dims = [chi if link_dim is None else link_dim[f] for f in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
        n = d * int(np.prod(dims)) if dims else d
        out.append(randn(rng, dtype, (d, *dims)) / np.sqrt(np.sqrt(n)))
    return out


def spin_ice_tensors(ga):
    """Indicator(i+j+k+l == 2) on every vertex of a 4-regular graph (test/test_beliefpropagation.jl:16-36)."""
    t = np.zeros((2, 2, 2, 2))
    for idx in np.ndindex(2, 2, 2, 2):
        if sum(idx) == 2:
            t[idx] = 1.0
    return [t.copy() for _ in range(ga.nv)]


def positive_messages(ga, link_dim, dtype, rng, mode="norm"):
    msgs = []
    for e in range(ga.ne):
        chi = link_dim[e]
        if mode == "norm":
            m = np.eye(chi) + 0.1 * np.abs(rng.standard_normal((chi, chi)))
            if np.dtype(dtype).kind == "c":
                m = m + 0.05j * rng.standard_normal((chi, chi))
        else:
            m = rng.random(chi) + 0.1
        msgs.append((m / m.sum()).astype(dtype))
    return msgs


def rel_err(got, want):
    return max(np.abs(np.asarray(g) - np.asarray(w)).max() / max(np.abs(np.asarray(w)).max(), 1e-300) for g, w in zip(got, want))
