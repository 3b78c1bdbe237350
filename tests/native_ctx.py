"""TEST INFRASTRUCTURE: a stand-in for `BPXContext` that runs the gate-application DEVICE CODE on the host through
tests/native/apply_host.cu (csrc/bpx_apply.cuh compiled for the host), so that the Python lowering of
itensornetworksnext.jl_b200/apply.py (names -> canonical layout, bond padding / slicing, layer batching) can be checked
against the reference's known answers without a GPU.  Only `-m "not gpu"` tests monkeypatch it in; the product always
talks to libbpx.so (CUDA)."""
import ctypes
import os
import subprocess

import numpy as np

P = ctypes.c_void_p
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "apply_host.cu")
HDR = os.path.join(HERE, "..", "itensornetworksnext.jl_b200", "csrc", "bpx_apply.cuh")
HDR2 = os.path.join(HERE, "..", "itensornetworksnext.jl_b200", "csrc", "bpx_apply2.cuh")
HDR3 = os.path.join(HERE, "..", "itensornetworksnext.jl_b200", "csrc", "bpx_expect2.cuh")
HDR4 = os.path.join(HERE, "..", "itensornetworksnext.jl_b200", "csrc", "bpx_apply3.cuh")
OUT = os.path.join(HERE, "native", "_build", "libapply_host.so")


def build_hostlib() -> ctypes.CDLL:
    """Compile tests/native/apply_host.cu (the device code of csrc/bpx_apply.cuh for the host) if stale; load it."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = not os.path.exists(OUT) or any(os.path.getmtime(f) > os.path.getmtime(OUT) for f in (SRC, HDR, HDR2, HDR3, HDR4))
    if stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-shared", "-o", OUT, SRC], check=True)
    return ctypes.CDLL(OUT)


def _ptr(a):
    return a.ctypes.data_as(P)


class HostHarnessContext:
    def __init__(self, hostlib, device=0, v2_block_rows=None):
        self.lib = hostlib
        self.calls = []  # (kind, number of gates) per device call -- lets tests see the batching
        self.v2_block_rows = v2_block_rows  # None: csrc/bpx_apply.cuh; an int: csrc/bpx_apply2.cuh with TSQR blocks of that many rows

    def set_graph(self, src, dst, slot, nv):
        self.src, self.dst, self.slot, self.nv = list(src), list(dst), list(slot), int(nv)
        self.ne = len(self.src)
        index = {(s, d): e for e, (s, d) in enumerate(zip(self.src, self.dst))}
        self.rev = [index[(d, s)] for s, d in zip(self.src, self.dst)]
        self.out = [[] for _ in range(self.nv)]
        for e, s in enumerate(self.src):
            self.out[s].append(e)
        for v in range(self.nv):
            self.out[v].sort(key=lambda e: self.slot[e])

    def set_dims(self, dtype, mode, phys_dim, link_dim):
        assert mode == "norm"
        self.dtype = np.dtype(dtype)
        self.phys, self.link_dim = list(phys_dim), [int(c) for c in link_dim]

    def set_site_tensors(self, tensors):
        self.sites = [np.asarray(t, dtype=self.dtype).ravel(order="F").copy() for t in tensors]

    def set_messages(self, msgs):
        self.msgs = [np.asarray(m, dtype=self.dtype).ravel(order="F").copy() for m in msgs]

    def _side(self, v):
        dims = np.array([self.link_dim[e] for e in self.out[v]], dtype=np.int32)
        incoming = [self.msgs[self.rev[e]] for e in self.out[v]]
        msgs = np.concatenate(incoming) if incoming else np.zeros(0, self.dtype)
        return len(self.out[v]), self.phys[v], dims, np.ascontiguousarray(msgs)

    def apply_two_site_gates(self, edges, ops, max_rank=0, normalize=False):
        self.calls.append(("two", len(edges)))
        code = 1 if self.dtype.kind == "c" else 0
        used, out = set(), []
        for e, op in zip(edges, ops):
            v1, v2, r = self.src[e], self.dst[e], self.rev[e]
            assert not ({v1, v2} & used), "gates of one call must be vertex-disjoint"
            used |= {v1, v2}
            z1, d1, dims1, m1 = self._side(v1)
            z2, d2, dims2, m2 = self._side(v2)
            chi = self.link_dim[e]
            o = np.asarray(op, dtype=self.dtype).ravel(order="F").copy()
            msg_out, sv = np.zeros(chi * chi, dtype=self.dtype), np.zeros(chi)
            args = (code, z1, d1, self.slot[e], _ptr(dims1), _ptr(self.sites[v1]), _ptr(m1), z2, d2, self.slot[r], _ptr(dims2),
                    _ptr(self.sites[v2]), _ptr(m2), _ptr(o), int(max_rank), int(bool(normalize)), _ptr(msg_out),
                    sv.ctypes.data_as(P))
            if self.v2_block_rows is None:
                rc = self.lib.apply_host_two_site(*args)
            else:
                need = 0
                for dims, slot, d in ((dims1, self.slot[e], d1), (dims2, self.slot[r], d2)):
                    rows = int(np.prod([dims[i] for i in range(len(dims)) if i != slot], dtype=np.int64))
                    cols = d * int(dims[slot])
                    need = max(need, 2 * rows, cols * cols + 2 * cols * (cols + self.v2_block_rows))
                rc = self.lib.apply_host_two_site_v2(*args, ctypes.c_int64(need))
            assert rc == 0
            self.msgs[e], self.msgs[r] = msg_out, msg_out.copy()
            out.append(sv)
        return out

    def apply_one_site_gates(self, vertices, ops, normalize=False):
        self.calls.append(("one", len(vertices)))
        code = 1 if self.dtype.kind == "c" else 0
        for v, op in zip(vertices, ops):
            z, d, dims, m = self._side(v)
            o = np.asarray(op, dtype=self.dtype).ravel(order="F").copy()
            self.lib.apply_host_one_site(code, z, d, _ptr(dims), _ptr(self.sites[v]), _ptr(m), _ptr(o), int(bool(normalize)))

    def edge_expect(self, edges, ops):
        """Two-site expectation values (csrc/bpx_expect2.cuh on the host): (numerators, denominators)."""
        self.calls.append(("expect2", len(edges)))
        code = 1 if self.dtype.kind == "c" else 0
        num, den = np.zeros(len(edges), dtype=self.dtype), np.zeros(len(edges), dtype=self.dtype)
        for i, (e, op) in enumerate(zip(edges, ops)):
            v1, v2, r = self.src[e], self.dst[e], self.rev[e]
            z1, d1, dims1, m1 = self._side(v1)
            z2, d2, dims2, m2 = self._side(v2)
            o = np.asarray(op, dtype=self.dtype).ravel(order="F").copy()
            a, b = np.zeros(1, dtype=self.dtype), np.zeros(1, dtype=self.dtype)
            rc = self.lib.apply_host_edge_expect(code, z1, d1, self.slot[e], _ptr(dims1), _ptr(self.sites[v1]), _ptr(m1), z2, d2,
                                                 self.slot[r], _ptr(dims2), _ptr(self.sites[v2]), _ptr(m2), _ptr(o), _ptr(a), _ptr(b))
            assert rc == 0
            num[i], den[i] = a[0], b[0]
        return num, den

    def get_site_tensor(self, v):
        return self.sites[v].copy()

    def close(self):
        pass

    # -- everything below is NOT device code: BP sweeps and beliefs come from the numpy oracle, so that host-side
    #    drivers that interleave gates and sweeps (ResidentState) can be exercised without a GPU --------------------
    def _oracle_problem(self):
        import __graft_entry__ as entry
        from itnn_b200.graphs import GraphArrays

        o = entry.import_oracle()
        row_ptr = [0]
        for v in range(self.nv):
            row_ptr.append(row_ptr[-1] + len(self.out[v]))
        ga = GraphArrays(vertices=list(range(self.nv)), vindex={v: v for v in range(self.nv)}, src=self.src, dst=self.dst,
                         rev=self.rev, slot=self.slot, row_ptr=row_ptr,
                         edge_index={(s, d): e for e, (s, d) in enumerate(zip(self.src, self.dst))})
        shapes = [(self.phys[v],) + tuple(self.link_dim[e] for e in self.out[v]) for v in range(self.nv)]
        tensors = [self.sites[v].reshape(shapes[v], order="F") for v in range(self.nv)]
        return o, o.make_problem(ga, tensors, "norm")

    def get_messages(self):
        return [m.reshape((c, c), order="F").copy() for m, c in zip(self.msgs, self.link_dim)]

    def _set(self, msgs):
        self.msgs = [np.asarray(m, dtype=self.dtype).ravel(order="F").copy() for m in msgs]

    def sweep(self, max_sweeps=1, tol=0.0, normalize=True):
        o, p = self._oracle_problem()
        msgs, res, done, self._history = self.get_messages(), float("inf"), 0, []
        for _ in range(max_sweeps):
            prev, msgs = msgs, o.sweep_jacobi(p, msgs, normalize)
            res, done = o.iterate_diff(msgs, prev), done + 1
            self._history.append(res)
            if tol > 0 and res < tol:
                break
        self._set(msgs)
        return res, done

    def sweep_sequence(self, edge_seq, max_sweeps=1, tol=0.0, normalize=True):
        o, p = self._oracle_problem()
        msgs, res, done, self._history = self.get_messages(), float("inf"), 0, []
        for _ in range(max_sweeps):
            prev = [m.copy() for m in msgs]
            msgs = o.sweep_sequential(p, msgs, list(edge_seq), normalize)
            res, done = o.iterate_diff(msgs, prev), done + 1
            self._history.append(res)
            if tol > 0 and res < tol:
                break
        self._set(msgs)
        return res, done

    def residual_history(self, n=4096):
        return np.array(getattr(self, "_history", []))

    def vertex_scalars(self):
        o, p = self._oracle_problem()
        return np.array(o.vertex_scalars(p, self.get_messages()))

    def edge_scalars(self):
        o, p = self._oracle_problem()
        return np.array(o.edge_scalars(p, self.get_messages()))

    def vertex_expect_numerators(self, ops):
        o, p = self._oracle_problem()
        msgs = self.get_messages()
        return np.array([o.vertex_scalar(p, msgs, v, np.asarray(ops[v])) for v in range(self.nv)])

    def iterate_diff(self, other):
        o, _ = self._oracle_problem()
        other = [np.asarray(m, dtype=self.dtype).reshape((c, c), order="F") for m, c in zip(other, self.link_dim)] \
            if not isinstance(other, np.ndarray) or other.ndim != 1 else \
            [other[sum(k * k for k in self.link_dim[:e]):][: c * c].reshape((c, c), order="F") for e, c in enumerate(self.link_dim)]
        return o.iterate_diff(self.get_messages(), other)

    def set_kernel_policy(self, kernel):
        pass
