"""ctypes binding of libbpx.so (include/bpx.h).  This is the Python twin of julia/BPX.jl's `ccall`s.

There is NO CPU fallback: if the shared library is missing the import of any compute path raises, and
`bpx_create` fails when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPX_LIB", os.path.join(_HERE, "csrc", "libbpx.so"))  # BPX_LIB: debug builds (tools/timing_*.py)
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "include", "bpx.h"))

BPX_OK = 0
BPX_F64, BPX_C64 = 0, 1
BPX_MODE_NORM, BPX_MODE_SINGLE = 0, 1
BPX_KERNEL_AUTO, BPX_KERNEL_GENERIC, BPX_KERNEL_ONCHIP, BPX_KERNEL_SLICED, BPX_KERNEL_VERTEX = 0, 1, 2, 3, 4
BPX_MAX_DEGREE = 12

_lib = None


class BPXError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libbpx error {status}: {message}")
        self.status = status


def header_symbols() -> list[str]:
    """Every function name declared in include/bpx.h (used by the CPU test that checks the exports)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpx_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load libbpx.so (building is `__graft_entry__.build()`'s job).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback for the BP hot path."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "bpx_version": (C.c_int, []),
        "bpx_create": (C.c_int, [C.c_int, P(vp)]),
        "bpx_create_multi": (C.c_int, [vp, C.c_int, P(vp)]),
        "bpx_num_devices": (C.c_int, [vp]),
        "bpx_set_owner": (C.c_int, [vp, vp]),
        "bpx_get_owner": (C.c_int, [vp, vp]),
        "bpx_destroy": (C.c_int, [vp]),
        "bpx_last_error": (C.c_char_p, [vp]),
        "bpx_set_graph": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "bpx_set_dims": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "bpx_num_vertices": (i64, [vp]),
        "bpx_num_edges": (i64, [vp]),
        "bpx_rev": (i64, [vp, i64]),
        "bpx_site_offset": (i64, [vp, i64]),
        "bpx_message_offset": (i64, [vp, i64]),
        "bpx_site_device_offset": (i64, [vp, i64]),
        "bpx_host_register": (C.c_int, [vp, vp, C.c_size_t]),
        "bpx_host_unregister": (C.c_int, [vp, vp]),
        "bpx_set_site_tensors": (C.c_int, [vp, vp]),
        "bpx_set_site_tensor": (C.c_int, [vp, i64, vp]),
        "bpx_set_messages": (C.c_int, [vp, vp]),
        "bpx_get_messages": (C.c_int, [vp, vp]),
        "bpx_get_message": (C.c_int, [vp, i64, vp]),
        "bpx_sweep": (C.c_int, [vp, C.c_int, dbl, C.c_int, P(dbl), P(C.c_int)]),
        "bpx_sweep_host": (C.c_int, [vp, vp, vp, C.c_int, P(dbl)]),
        "bpx_sweep_async": (C.c_int, [vp, C.c_int, C.c_int]),
        "bpx_set_profiling": (C.c_int, [vp, C.c_int]),
        "bpx_bucket_time": (C.c_int, [vp, C.c_int, P(dbl), P(i64)]),
        "bpx_sweep_sequence": (C.c_int, [vp, vp, i64, C.c_int, dbl, C.c_int, P(dbl), P(C.c_int)]),
        "bpx_residual_history": (C.c_int, [vp, vp, C.c_int, P(C.c_int)]),
        "bpx_last_residual": (C.c_int, [vp, P(dbl)]),
        "bpx_iterate_diff": (C.c_int, [vp, vp, P(dbl)]),
        "bpx_vertex_scalars": (C.c_int, [vp, vp]),
        "bpx_edge_scalars": (C.c_int, [vp, vp]),
        "bpx_apply_stats": (C.c_int, [vp, C.POINTER(C.c_int64), C.c_int]),
        "bpx_bethe_free_energy": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
        "bpx_bethe_free_energy_parts": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "bpx_vertex_expect_numerators": (C.c_int, [vp, vp, vp]),
        "bpx_apply_two_site_gates": (C.c_int, [vp, i64, vp, vp, C.c_int, C.c_int, vp]),
        "bpx_apply_one_site_gates": (C.c_int, [vp, i64, vp, vp, C.c_int]),
        "bpx_get_site_tensor": (C.c_int, [vp, i64, vp]),
        "bpx_edge_expect": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "bpx_num_buckets": (C.c_int, [vp]),
        "bpx_bucket_info": (C.c_int, [vp, C.c_int, P(i64)]),
        "bpx_set_kernel_policy": (C.c_int, [vp, C.c_int]),
        "bpx_counters": (C.c_int, [vp, P(i64), C.c_int]),
        "bpx_set_stream": (C.c_int, [vp, vp]),
        "bpx_device_messages": (vp, [vp]),
        "bpx_device_site_tensors": (vp, [vp]),
        "bpx_synchronize": (C.c_int, [vp]),
        "bpx_set_partition": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "bpx_halo_export": (C.c_int, [vp, vp]),
        "bpx_halo_connect": (C.c_int, [vp, C.c_int, vp]),
        "bpx_num_cut_edges": (i64, [vp]),
        "bpx_peer_barrier": (C.c_int, [vp]),
        "bpx_fill_synthetic": (C.c_int, [vp, C.c_uint64]),
        "bpx_fill_randn": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._bpx_signatures = sig
    _lib = lib
    return lib
