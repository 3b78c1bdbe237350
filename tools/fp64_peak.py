"""FP64 roofline denominator: cuBLAS DGEMM 8192^3 through torch.matmul(float64), measured the way the
driver measured the bf16 entry of MEASURED_PEAKS.json -- best of 10 (burst) and back to back for 4 s (sustained), with
the clocks seen.  Writes one JSON object (profiles/fp64_peak.json is a committed copy of one run on this pool's B200).

    python tools/fp64_peak.py [out.json]
"""
import json
import subprocess
import sys
import time

import torch


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,"
                              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "power_w": float(out[2]), "sw_power_cap": out[3].strip(),
                "hw_slowdown": out[4].strip(), "sw_thermal_slowdown": out[5].strip()}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def measure(n=8192, burst_reps=10, sustained_s=4.0):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = float("inf")
    for _ in range(burst_reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
        time.sleep(0.05)
    burst = flop / (best * 1e-3) * 1e-12
    # sustained: back to back for >= sustained_s seconds, one event pair around the lot
    per = best * 1e-3
    reps = max(4, int(sustained_s / per))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mid = None
    for i in range(reps):
        torch.matmul(a, b, out=c)
        if i == reps // 2:
            mid = True
    e1.record()
    time.sleep(min(2.0, 0.5 * reps * per))
    under_load = clocks()
    e1.synchronize()
    sustained = flop * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    return {"fp64_tflops": burst, "fp64_tflops_sustained": sustained, "n": n, "burst_reps": burst_reps, "sustained_reps": reps,
            "sustained_seconds": e0.elapsed_time(e1) * 1e-3, "clocks_under_load": under_load,
            "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
            "how": "torch.matmul float64 8192^3 (2*N^3): best of 10 with 50 ms pauses (burst) and back to back for ~4 s (sustained), CUDA events"}


if __name__ == "__main__":
    r = measure()
    s = json.dumps(r, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s + "\n")
