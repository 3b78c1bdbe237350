"""numpy restatement of the shared counter-based RNG (libbpx `bpx_fill_randn`: splitmix64 counter -> Box-Muller), so that
the reference arm of bench.py can build the synthetic inputs WITHOUT loading the product library.  TEST INFRASTRUCTURE
ONLY (tests/test_host_api.py pins it against the library bit for bit, up to libm's last ulp in log/cos)."""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def _randn(seed: int, stream: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        s = _splitmix64(np.array([np.uint64(stream) + np.uint64(0x632BE59BD9B4E019)], dtype=np.uint64))
        base = _splitmix64(np.array([np.uint64(seed)], dtype=np.uint64) ^ s)[0]
        i = np.arange(n, dtype=np.uint64)
        a = _splitmix64(base + np.uint64(2) * i)
        b = _splitmix64(base + np.uint64(2) * i + np.uint64(1))
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925286766559 * u2)


def fill_randn(seed: int, stream: int, dtype, n: int) -> np.ndarray:
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        r = _randn(seed, stream, 2 * n) * 0.70710678118654752440
        return (r[0::2] + 1j * r[1::2]).astype(dtype)
    return _randn(seed, stream, n).astype(dtype)
