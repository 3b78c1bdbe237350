"""`BPXContext`: one libbpx context (one GPU) driven with numpy host buffers.

Thin, mechanical wrapper over the C ABI (include/bpx.h); all arithmetic happens in the CUDA kernels.
The canonical layout is column-major: site tensors and messages are handed over as Fortran-order flats.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BPXError


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def cast_to(a, dtype, what: str = "array") -> np.ndarray:
    """`np.asarray(a, dtype)` that refuses to throw an imaginary part away: a complex operator / message / tensor handed to
    a Float64 context is an error (the reference promotes to ComplexF64, ITensorBase.apply), not a silent truncation."""
    a = np.asarray(a)
    dtype = np.dtype(dtype)
    if a.dtype.kind == "c" and dtype.kind != "c" and a.size and np.any(a.imag != 0):
        raise TypeError(f"complex {what} for a Float64 context would lose its imaginary part: promote the state / iterate "
                        "to ComplexF64 first")
    return a.real.astype(dtype) if (a.dtype.kind == "c" and dtype.kind != "c") else a.astype(dtype, copy=False)


def pack_operators(ops: Sequence[np.ndarray], dtype) -> np.ndarray:
    """The operators of a batch, each flattened column-major, back to back.  Layers of a circuit are thousands of small
    operators of one shape: those are stacked and transposed in one go instead of one `ravel` per operator (4 x faster on
    a 2 000-gate layer, where the per-operator loop cost as much host time as a quarter of the kernels)."""
    if not len(ops):
        return np.empty(0, np.dtype(dtype))
    first = np.asarray(ops[0])
    if all(isinstance(o, np.ndarray) and o.shape == first.shape and o.dtype == first.dtype for o in ops):
        a = cast_to(np.stack(ops), dtype, "operator")
        return np.ascontiguousarray(a.transpose((0,) + tuple(range(a.ndim - 1, 0, -1)))).reshape(-1)
    return np.ascontiguousarray(np.concatenate([cast_to(o, dtype, "operator").ravel(order="F") for o in ops]))


class BPXContext:
    """`BPXContext(0)`: one device.  `BPXContext(devices=[0, 1, ...])`: ONE context over several devices of this process
    (bpx_create_multi): same methods, the library partitions the vertices and exchanges cut-edge messages over NVLink."""

    def __init__(self, device: int = 0, devices: Optional[Sequence[int]] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if devices is not None:
            devs = np.ascontiguousarray(list(devices), dtype=np.int32)
            rc = self.lib.bpx_create_multi(_ptr(devs), len(devs), C.byref(h))
            device = int(devs[0]) if len(devs) else 0
        else:
            rc = self.lib.bpx_create(int(device), C.byref(h))
        if rc != 0:
            raise BPXError(rc, self.lib.bpx_last_error(None).decode())
        self.h = h
        self.device = device
        self.devices = [int(d) for d in devices] if devices is not None else [int(device)]
        self.dtype = None
        self.mode = None
        self.nv = self.ne = 0

    # -- plumbing ----------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise BPXError(rc, self.lib.bpx_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bpx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- problem description -----------------------------------------------------------------------
    def set_graph(self, src: Sequence[int], dst: Sequence[int], slot: Sequence[int], nv: int):
        src = np.ascontiguousarray(src, dtype=np.int64)
        dst = np.ascontiguousarray(dst, dtype=np.int64)
        slot = np.ascontiguousarray(slot, dtype=np.int32)
        self._check(self.lib.bpx_set_graph(self.h, int(nv), len(src), _ptr(src), _ptr(dst), _ptr(slot)))
        self.nv, self.ne = int(nv), len(src)
        self._src = src

    def set_dims(self, dtype, mode: str, phys_dim: Optional[Sequence[int]], link_dim: Sequence[int]):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.complex128)):
            raise TypeError("libbpx computes in Float64 / ComplexF64 only")
        self.mode = mode
        code = _lib.BPX_F64 if self.dtype == np.float64 else _lib.BPX_C64
        m = {"norm": _lib.BPX_MODE_NORM, "single": _lib.BPX_MODE_SINGLE}[mode]
        pd = None if phys_dim is None else np.ascontiguousarray(phys_dim, dtype=np.int32)
        ld = np.ascontiguousarray(link_dim, dtype=np.int32)
        self._check(self.lib.bpx_set_dims(self.h, code, m, _ptr(pd), _ptr(ld)))
        # packed host layouts (the library's bpx_site_offset / bpx_message_offset, computed here in bulk: one ctypes call
        # per vertex and edge takes seconds on lattices of millions of vertices); the totals are checked against the library
        n_site = np.ones(self.nv, dtype=np.int64) if pd is None or mode == "single" else pd.astype(np.int64)
        np.multiply.at(n_site, self._src, ld.astype(np.int64))
        self.site_off = np.concatenate([[0], np.cumsum(n_site)]).astype(np.int64)
        n_msg = ld.astype(np.int64) ** 2 if mode == "norm" else ld.astype(np.int64)
        self.msg_off = np.concatenate([[0], np.cumsum(n_msg)]).astype(np.int64)
        if (self.site_off[-1] != self.lib.bpx_site_offset(self.h, self.nv) or self.msg_off[-1] != self.lib.bpx_message_offset(self.h, self.ne)
                or (self.nv and self.site_off[self.nv // 2] != self.lib.bpx_site_offset(self.h, self.nv // 2))
                or (self.ne and self.msg_off[self.ne // 2] != self.lib.bpx_message_offset(self.h, self.ne // 2))):
            raise RuntimeError("packed layout mismatch between the host mirror and libbpx")
        self.link_dim = ld

    def set_owner(self, owner: Sequence[int]):
        """Multi-device contexts: owner[v] = index of the device that updates the out-edges of v (after set_dims)."""
        own = np.ascontiguousarray(owner, dtype=np.int32)
        self._check(self.lib.bpx_set_owner(self.h, _ptr(own)))

    def get_owner(self) -> np.ndarray:
        out = np.zeros(self.nv, dtype=np.int32)
        self._check(self.lib.bpx_get_owner(self.h, _ptr(out)))
        return out

    # -- data --------------------------------------------------------------------------------------
    def pack_sites(self, tensors: Sequence[np.ndarray]) -> np.ndarray:
        out = np.empty(int(self.site_off[-1]), dtype=self.dtype)
        for v, t in enumerate(tensors):
            n = int(self.site_off[v + 1] - self.site_off[v])
            if t.size != n:
                raise ValueError(f"site tensor {v} has {t.size} elements, expected {n}")
            out[self.site_off[v]:self.site_off[v + 1]] = cast_to(t, self.dtype, "site tensor").ravel(order="F")
        return out

    def pack_messages(self, msgs: Sequence[np.ndarray]) -> np.ndarray:
        if isinstance(msgs, np.ndarray) and msgs.ndim == 1:  # already packed
            if msgs.size != int(self.msg_off[-1]):
                raise ValueError(f"packed messages have {msgs.size} elements, expected {int(self.msg_off[-1])}")
            return np.ascontiguousarray(cast_to(msgs, self.dtype, "message set"))
        out = np.empty(int(self.msg_off[-1]), dtype=self.dtype)
        for e, m in enumerate(msgs):
            n = int(self.msg_off[e + 1] - self.msg_off[e])
            if m.size != n:
                raise ValueError(f"message {e} has {m.size} elements, expected {n}")
            out[self.msg_off[e]:self.msg_off[e + 1]] = cast_to(m, self.dtype, "message").ravel(order="F")
        return out

    def unpack_messages(self, flat: np.ndarray) -> List[np.ndarray]:
        out = []
        for e in range(self.ne):
            chi = int(self.link_dim[e])
            shape = (chi, chi) if self.mode == "norm" else (chi,)
            out.append(flat[self.msg_off[e]:self.msg_off[e + 1]].reshape(shape, order="F").copy())
        return out

    def set_site_tensors(self, tensors):
        flat = tensors if isinstance(tensors, np.ndarray) and tensors.ndim == 1 else self.pack_sites(tensors)
        flat = np.ascontiguousarray(flat, dtype=self.dtype)
        self._check(self.lib.bpx_set_site_tensors(self.h, _ptr(flat)))

    def set_messages(self, msgs):
        flat = msgs if isinstance(msgs, np.ndarray) and msgs.ndim == 1 else self.pack_messages(msgs)
        flat = np.ascontiguousarray(flat, dtype=self.dtype)
        self._check(self.lib.bpx_set_messages(self.h, _ptr(flat)))

    def get_messages_flat(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(int(self.msg_off[-1]), dtype=self.dtype)
        self._check(self.lib.bpx_get_messages(self.h, _ptr(out)))
        return out

    def get_messages(self) -> List[np.ndarray]:
        return self.unpack_messages(self.get_messages_flat())

    def fill_synthetic(self, seed: int = 123):
        """Generate the synthetic benchmark inputs on the device (no host staging)."""
        self._check(self.lib.bpx_fill_synthetic(self.h, int(seed)))

    def get_site_tensor(self, v: int) -> np.ndarray:
        """Download one canonical site tensor (flat, column-major): `state[v]` after gates were applied."""
        if len(self.devices) == 1 and self.lib.bpx_site_device_offset(self.h, int(v)) < 0:
            raise KeyError(f"site tensor {v} is not resident on this rank")
        out = np.empty(int(self.site_off[v + 1] - self.site_off[v]), dtype=self.dtype)
        self._check(self.lib.bpx_get_site_tensor(self.h, int(v), _ptr(out)))
        return out

    # -- gate application (src/apply/apply_operators.jl:213-283) ------------------------------------
    def apply_two_site_gates(self, edges: Sequence[int], ops: Sequence[np.ndarray], max_rank: int = 0,
                             normalize: bool = False) -> List[np.ndarray]:
        """A batch of vertex-disjoint two-site gates, in place on the device.  ops[g][o1, o2, i1, i2] with 1 = src and
        2 = dst of directed edge edges[g].  Returns the kept singular values per gate (zero-padded to the link dim)."""
        e = np.ascontiguousarray(edges, dtype=np.int64)
        flat = pack_operators(ops, self.dtype)
        dims = [int(self.link_dim[i]) if 0 <= i < self.ne else 0 for i in e]  # bad ids are reported by the library
        sv = np.zeros(max(1, sum(dims)), dtype=np.float64)
        self._check(self.lib.bpx_apply_two_site_gates(self.h, len(e), _ptr(e), _ptr(flat), int(max_rank),
                                                      int(bool(normalize)), _ptr(sv)))
        if dims and min(dims) == max(dims) and dims[0] > 0:
            return list(sv[:len(dims) * dims[0]].reshape(len(dims), dims[0]))  # rows of one fresh array
        out, o = [], 0
        for c in dims:
            out.append(sv[o:o + c].copy())
            o += c
        return out

    def apply_stats(self, reset: bool = False):
        """(two-site gates applied by the Gram-path kernel, gates it declined and the step-by-step kernel applied)."""
        out = (C.c_int64 * 2)()
        self._check(self.lib.bpx_apply_stats(self.h, out, int(bool(reset))))
        return int(out[0]), int(out[1])

    def apply_one_site_gates(self, vertices: Sequence[int], ops: Sequence[np.ndarray], normalize: bool = False):
        v = np.ascontiguousarray(vertices, dtype=np.int64)
        flat = pack_operators(ops, self.dtype)
        self._check(self.lib.bpx_apply_one_site_gates(self.h, len(v), _ptr(v), _ptr(flat), int(bool(normalize))))

    def edge_expect(self, edges: Sequence[int], ops: Sequence[np.ndarray]):
        """Two-site expectation values in the BP environment: (numerators, denominators) per listed directed edge;
        ops[g][o1, o2, i1, i2] with 1 = src and 2 = dst of edges[g].  Read-only (edges may share vertices)."""
        e = np.ascontiguousarray(edges, dtype=np.int64)
        flat = pack_operators(ops, self.dtype)
        num, den = np.zeros(max(1, len(e)), dtype=self.dtype), np.zeros(max(1, len(e)), dtype=self.dtype)
        self._check(self.lib.bpx_edge_expect(self.h, len(e), _ptr(e), _ptr(flat), _ptr(num), _ptr(den)))
        return num[:len(e)], den[:len(e)]

    # -- hot path ----------------------------------------------------------------------------------
    def sweep(self, max_sweeps: int = 1, tol: float = 0.0, normalize: bool = True):
        res, done = C.c_double(), C.c_int()
        self._check(self.lib.bpx_sweep(self.h, int(max_sweeps), float(tol), int(bool(normalize)), C.byref(res), C.byref(done)))
        return res.value, done.value

    def sweep_host(self, flat_in: np.ndarray, flat_out: np.ndarray, normalize: bool = True) -> float:
        """Upload messages, one synchronous sweep, download messages + residual (one host sync)."""
        res = C.c_double()
        self._check(self.lib.bpx_sweep_host(self.h, _ptr(flat_in), _ptr(flat_out), int(bool(normalize)), C.byref(res)))
        return res.value

    def host_register(self, arr: np.ndarray):
        """Page-lock `arr` so that `sweep_host` can stream through it (cudaHostRegister)."""
        self._check(self.lib.bpx_host_register(self.h, _ptr(arr), C.c_size_t(arr.nbytes)))

    def host_unregister(self, arr: np.ndarray):
        self._check(self.lib.bpx_host_unregister(self.h, _ptr(arr)))

    def sweep_async(self, n_sweeps: int = 1, normalize: bool = True):
        self._check(self.lib.bpx_sweep_async(self.h, int(n_sweeps), int(bool(normalize))))

    def peer_barrier(self):
        """Device-side barrier over all connected ranks, enqueued on the context's stream (collective)."""
        self._check(self.lib.bpx_peer_barrier(self.h))

    def set_profiling(self, enable: bool):
        self._check(self.lib.bpx_set_profiling(self.h, int(bool(enable))))

    def bucket_time(self, bucket: int):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.bpx_bucket_time(self.h, int(bucket), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_residual(self) -> float:
        """Residual of the last enqueued sweep (synchronises)."""
        res = C.c_double()
        self._check(self.lib.bpx_last_residual(self.h, C.byref(res)))
        return res.value

    def sweep_sequence(self, edge_seq: Sequence[int], max_sweeps: int = 1, tol: float = 0.0, normalize: bool = True):
        seq = np.ascontiguousarray(edge_seq, dtype=np.int64)
        res, done = C.c_double(), C.c_int()
        self._check(self.lib.bpx_sweep_sequence(self.h, _ptr(seq), len(seq), int(max_sweeps), float(tol),
                                                int(bool(normalize)), C.byref(res), C.byref(done)))
        return res.value, done.value

    def residual_history(self, n: int = 4096) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        k = C.c_int()
        self._check(self.lib.bpx_residual_history(self.h, _ptr(out), n, C.byref(k)))
        return out[:k.value].copy()

    def iterate_diff(self, other) -> float:
        flat = other if isinstance(other, np.ndarray) and other.ndim == 1 else self.pack_messages(other)
        flat = np.ascontiguousarray(flat, dtype=self.dtype)
        res = C.c_double()
        self._check(self.lib.bpx_iterate_diff(self.h, _ptr(flat), C.byref(res)))
        return res.value

    # -- beliefs -----------------------------------------------------------------------------------
    def vertex_scalars(self) -> np.ndarray:
        out = np.empty(self.nv, dtype=self.dtype)
        self._check(self.lib.bpx_vertex_scalars(self.h, _ptr(out)))
        return out

    def edge_scalars(self) -> np.ndarray:
        out = np.empty(self.ne // 2, dtype=self.dtype)
        self._check(self.lib.bpx_edge_scalars(self.h, _ptr(out)))
        return out

    def bethe_free_energy(self):
        """messagecache.jl:185-201 reduced on the device (bpx_bethe_free_energy): a float, or a complex number when the
        reference's promotion rule applies."""
        out = (C.c_double * 2)()
        promoted = C.c_int()
        self._check(self.lib.bpx_bethe_free_energy(self.h, out, C.byref(promoted)))
        return complex(out[0], out[1]) if promoted.value else float(out[0])

    def bethe_free_energy_parts(self) -> np.ndarray:
        out = (C.c_double * 7)()
        self._check(self.lib.bpx_bethe_free_energy_parts(self.h, out))
        return np.array(out[:])

    def vertex_expect_numerators(self, ops: Sequence[np.ndarray]) -> np.ndarray:
        flat = pack_operators(ops, self.dtype)
        out = np.empty(self.nv, dtype=self.dtype)
        self._check(self.lib.bpx_vertex_expect_numerators(self.h, _ptr(flat), _ptr(out)))
        return out

    # -- introspection -----------------------------------------------------------------------------
    def buckets(self):
        out = []
        for b in range(self.lib.bpx_num_buckets(self.h)):
            info = (C.c_int64 * 8)()
            self._check(self.lib.bpx_bucket_info(self.h, b, info))
            out.append(dict(degree=info[0], chi=info[1], phys=info[2], vertices=info[3], edges=info[4], kernel=info[5],
                            leader=info[6]))
        return out

    def set_kernel_policy(self, kernel: int):
        self._check(self.lib.bpx_set_kernel_policy(self.h, int(kernel)))

    def counters(self, reset: bool = False):
        out = (C.c_int64 * 3)()
        self._check(self.lib.bpx_counters(self.h, out, int(reset)))
        return dict(launches=out[0], updates=out[1], sweeps=out[2])

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.bpx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.bpx_synchronize(self.h))


def fill_randn(seed: int, stream: int, dtype, n: int) -> np.ndarray:
    """Shared deterministic RNG of the library (host function; needs no GPU)."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    out = np.empty(n, dtype=dtype)
    code = _lib.BPX_F64 if dtype == np.float64 else _lib.BPX_C64
    rc = lib.bpx_fill_randn(seed, stream, code, n, _ptr(out))
    if rc != 0:
        raise BPXError(rc, "bpx_fill_randn: invalid arguments")
    return out
