"""ctypes wrapper of oracle/liboracle_bp.so (the plain-C restatement, bp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_bp.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "bp_oracle.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.oracle_iterate_diff.restype = C.c_double
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracle:
    """Holds the canonical flat arrays of one problem (norm mode) and runs sweeps in C."""

    def __init__(self, ga, phys_dim: Sequence[int], link_dim: Sequence[int], tensors: Sequence[np.ndarray], dtype):
        self.lib = load()
        self.dtype = np.dtype(dtype)
        self.is_complex = int(self.dtype.kind == "c")
        self.nv, self.ne = ga.nv, ga.ne
        self.src = np.ascontiguousarray(ga.src, dtype=np.int64)
        self.rev = np.ascontiguousarray(ga.rev, dtype=np.int64)
        self.slot = np.ascontiguousarray(ga.slot, dtype=np.int64)
        self.row_ptr = np.ascontiguousarray(ga.row_ptr, dtype=np.int64)
        self.phys_dim = np.ascontiguousarray(phys_dim, dtype=np.int32)
        self.link_dim = np.ascontiguousarray(link_dim, dtype=np.int32)
        sizes = [int(np.prod(t.shape)) for t in tensors]
        self.site_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.msg_off = np.concatenate([[0], np.cumsum(self.link_dim.astype(np.int64) ** 2)]).astype(np.int64)
        self.sites = np.concatenate([np.asarray(t, dtype=self.dtype).ravel(order="F") for t in tensors])

    def pack(self, msgs: Sequence[np.ndarray]) -> np.ndarray:
        return np.concatenate([np.asarray(m, dtype=self.dtype).ravel(order="F") for m in msgs])

    def unpack(self, flat: np.ndarray) -> List[np.ndarray]:
        return [flat[self.msg_off[e]:self.msg_off[e + 1]].reshape((self.link_dim[e],) * 2, order="F").copy()
                for e in range(self.ne)]

    def sweep_jacobi(self, flat_in: np.ndarray, normalize: bool = True, variant: int = 0, nthreads: int = 0,
                     edges: Optional[Sequence[int]] = None) -> np.ndarray:
        out = np.empty_like(flat_in)
        el = None if edges is None else np.ascontiguousarray(edges, dtype=np.int64)
        rc = self.lib.oracle_sweep_jacobi(
            self.is_complex, C.c_int64(self.nv), C.c_int64(self.ne), _p(self.src), _p(self.rev), _p(self.slot),
            _p(self.row_ptr), _p(self.phys_dim), _p(self.link_dim), _p(self.site_off), _p(self.msg_off), _p(self.sites),
            _p(flat_in), _p(out), _p(el), C.c_int64(0 if el is None else len(el)), int(normalize), int(variant), int(nthreads))
        if rc:
            raise RuntimeError(f"oracle_sweep_jacobi failed ({rc})")
        return out

    def sweep_sequential(self, flat: np.ndarray, edge_seq: Sequence[int], normalize: bool = True, variant: int = 0) -> np.ndarray:
        out = flat.copy()
        seq = np.ascontiguousarray(edge_seq, dtype=np.int64)
        rc = self.lib.oracle_sweep_sequential(
            self.is_complex, C.c_int64(self.nv), C.c_int64(self.ne), _p(self.src), _p(self.rev), _p(self.slot),
            _p(self.row_ptr), _p(self.phys_dim), _p(self.link_dim), _p(self.site_off), _p(self.msg_off), _p(self.sites),
            _p(out), _p(seq), C.c_int64(len(seq)), int(normalize), int(variant))
        if rc:
            raise RuntimeError(f"oracle_sweep_sequential failed ({rc})")
        return out

    def iterate_diff(self, a: np.ndarray, b: np.ndarray) -> float:
        return float(self.lib.oracle_iterate_diff(self.is_complex, C.c_int64(self.ne), _p(self.msg_off), _p(a), _p(b)))

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())
