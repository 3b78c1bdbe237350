"""Inter-warp synchronisation of the gate-application kernels (csrc/bpx_apply.cuh, csrc/bpx_apply2.cuh, the phases of
csrc/bpx_apply3.cuh on warps of one lane) checked WITHOUT a
GPU: tests/native/apply_race_check.cu runs the same `__host__ __device__` code with one thread per warp -- or per LANE,
with warp barriers standing in for `__syncwarp()` and the shuffle reductions -- and a real barrier behind `Team::sync()`,
under ThreadSanitizer.  A missing barrier between two phases executed by different warps is a
data-race report (or a result that differs from the sequential schedule); with the barriers removed TSAN reports
thousands of races on this very program, so the check has teeth."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "apply_race_check.cu")
EXE = os.path.join(HERE, "native", "_build", "apply_race_check")
HDRS = [os.path.join(HERE, "..", "itensornetworksnext.jl_b200", "csrc", f)
        for f in ("bpx_apply.cuh", "bpx_apply2.cuh", "bpx_apply3.cuh", "bpx_expect2.cuh", "bpx_common.cuh")]


def test_gate_kernels_are_race_free_across_warps():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in [SRC] + HDRS):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-O1", "-g", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fsanitize=thread",
                        "-Xcompiler", "-pthread", "-o", EXE, SRC], check=True, capture_output=True)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    if "FATAL: ThreadSanitizer" in out:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + out.strip().splitlines()[0])
    assert "ThreadSanitizer: data race" not in out, out[-3000:]
    assert r.returncode == 0 and "all schedules agree" in out, out[-3000:]
    # (8 gate problems + 2 expectation problems) x 5 schedules + 7 Gram-path problems x 3 schedules
    assert out.count(" ok") >= 71 and "MISMATCH" not in out
