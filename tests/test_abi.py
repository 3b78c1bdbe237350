"""The C-ABI library loads and exports every symbol include/bpx.h declares (no compute without a GPU)."""
import ctypes as C

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    names = pkg._lib.header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # every declared function also has a ctypes signature in the binding
    assert set(names) <= set(lib._bpx_signatures)
    assert lib.bpx_version() == 100


def test_create_fails_loudly_without_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.BPXError) as ei:
        pkg.BPXContext(0)
    assert "no CPU fallback" in str(ei.value)


def test_shared_rng_is_deterministic_and_normal(pkg):
    a = pkg.fill_randn(123, 7, np.float64, 100000)
    b = pkg.fill_randn(123, 7, np.float64, 100000)
    c = pkg.fill_randn(123, 8, np.float64, 100000)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1) < 0.02
    z = pkg.fill_randn(123, 7, np.complex128, 50000)
    assert abs((abs(z) ** 2).mean() - 1) < 0.03
    # prefix property: out[i] depends only on (seed, stream, i)
    assert np.array_equal(pkg.fill_randn(123, 7, np.float64, 10), a[:10])


def _build_demo(tmp_path):
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "itensornetworksnext.jl_b200", "csrc")
    exe = str(tmp_path / "bpx_demo")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "bpx_demo.c"), "-o", exe, "-L" + libdir, "-lbpx", "-Wl,-rpath," + libdir, "-lm"],
                   check=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_header_is_plain_c_and_a_c_client_links(pkg, tmp_path):
    """include/bpx.h compiles as strict C99 (no C++ or torch types at the boundary); a C client links against libbpx.so
    and, without a GPU, is told loudly that there is no CPU fallback."""
    import torch

    pkg._lib.load()
    r = _build_demo(tmp_path)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
