#!/usr/bin/env python
"""Benchmark of the BP message-update hot path (BASELINE.json metric: BP message updates/s).

  python bench.py --gpus N --steps K --warmup W [--workload cfg2] [--impl reference]

A "step" is ONE synchronous BP sweep (every directed edge updated once) over a synthetic PEPS norm
network.  N = 1 runs BASELINE config 2 (32x32 square lattice, chi = 8, d = 2, Float64); for N > 1 the
lattice is vertex-partitioned into N strips of 32 rows (weak scaling: every rank owns a 32x32 block of a
(32N)x32 lattice), cut-edge messages are pushed to the neighbour ranks over NVLink every sweep and the
residual is max-reduced over the ranks.

One JSON line on rank 0 (see the task contract): `value` = updates/s with everything resident in HBM,
`e2e` = the same through the C ABI with HOST buffers (message upload + sweep + message download per
step), `roofline` for the dominant bucket's kernel, `cpu_baseline` = the CPU restatement of the reference
algorithm (oracle/bp_oracle.c, all host cores) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

# DRAM traffic per launch of the dominant kernel, from one `ncu --set full` capture each
# (dram__bytes_read.sum + dram__bytes_write.sum; summaries committed under profiles/).  Keyed by (workload, n_gpus).
PROFILED_TRAFFIC = {
    ("cfg2", 1): (62.28e6 + 0.59e6, "profiles/r1d_onchip_c8_final_ncu_summary.csv"),
    ("cfg4", 1): (272.88e6 + 9.31e6, "profiles/r1e_onchip_c16_cfg4_ncu_summary.csv"),
    ("cfg3", 1): (6.72e6 + 0.0, "profiles/r1f_onchip_c16c_cfg3_ncu_summary.csv (the 1.2 MB of new messages stay in L2)"),
    ("ising", 1): (230.85e6 + 48.74e6, "profiles/r1k_vertex_ising_v3_ncu_summary.csv (each message is read twice, as input and as the old "
                   "value, but comes from DRAM once; + 29 MB of descriptors; part of the output stays in L2)"),
}
# sliced chi=16 kernel: 926.9 MB read + 415.0 MB written for 196 degree-4 vertices (profiles/r1c_sliced_c16_ncu_summary.csv)
SLICED_TRAFFIC_PER_VERTEX = (926.93e6 + 415.05e6) / 196.0

METRIC = "bp_message_updates_per_s"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "cfg5s", "cfg2c", "ising"])
    ap.add_argument("--kernel", type=int, default=0, help="force a kernel family (include/bpx.h BPX_KERNEL_*)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--converge", type=float, default=0.0, metavar="TOL",
                    help="additionally run BP to convergence (maxiter 200, StopWhenConverged(TOL)) from the initial messages and report "
                         "it under \"convergence\" (untimed by the step metric)")
    ap.add_argument("--dump-timing", action="store_true", help="debug (BPX_ONCHIP_TIMING builds): per-CTA globaltimer stamps of the last sweep")
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between timed steps: 256 MiB memset, optionally followed by a read pass over the same buffer "
                         "(leaves L2 full of clean instead of dirty lines)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def build_workload(name: str, world: int):
    """-> (SyntheticProblem, owner list or None, description)."""
    from itnn_b200 import graphs, problems

    if name == "ising":  # not a BASELINE config: the HBM-bound single-layer bucket of the path
        if world > 1:
            raise SystemExit("--workload ising is a single-GPU workload")
        return (problems.make_config("ising"), None,
                "1024x1024 periodic square-lattice Ising partition-function network (ising_network recipe, beta=0.3), single layer, chi=2, Float64")
    if name == "cfg5s":  # cfg5's buckets on a lattice the default run can generate quickly
        g = graphs.named_grid((24, 24))
        p = problems.make_config("cfg5", graph=g)
        return p, None, "24x24 square-lattice PEPS norm network (cfg5 buckets), chi=16, d=2, Float64"
    if world == 1 or name != "cfg2":
        if name == "cfg5" and world > 1:  # the north-star target: 256x256 vertex-sharded into strips of rows
            p = problems.make_config(name, host_data=False)
            owner = [(v[1] - 1) * world // 256 for v in p.ga.vertices]
            return p, owner, f"256x256 square-lattice PEPS norm network, chi=16, d=2, Float64, {world} strips of {256 // world} rows (STRONG scaling)"
        if name == "cfg4" and world > 1:  # BASELINE config 4 "at 1/2/4/8 B200": slabs of the periodic cube
            p = problems.make_config(name)
            owner = [(v[2] - 1) * world // 16 for v in p.ga.vertices]
            return p, owner, f"16x16x16 periodic cubic PEPS norm network, chi=4, d=2, Float64, {world} slabs of {16 // world} planes (STRONG scaling)"
        p = problems.make_config(name, host_data=(name != "cfg5"))
        desc = {
            "cfg1": "4x4 square-lattice PEPS norm network, chi=2, d=2, Float64",
            "cfg2": "32x32 square-lattice PEPS norm network, chi=8, d=2, Float64",
            "cfg3": "heavy-hex 127-site PEPS norm network, chi=16, d=2, ComplexF64",
            "cfg4": "16x16x16 periodic cubic PEPS norm network, chi=4, d=2, Float64",
            "cfg5": "256x256 square-lattice PEPS norm network, chi=16, d=2, Float64",
            "cfg2c": "32x32 square-lattice PEPS norm network, chi=8, d=2, ComplexF64 (complex twin of config 2; not a BASELINE config)",
        }[name]
        return p, None, desc
    g = graphs.named_grid((32, 32 * world))
    p = problems.make_config("cfg2", graph=g)
    owner = [(v[1] - 1) // 32 for v in p.ga.vertices]
    return p, owner, f"32x{32 * world} square-lattice PEPS norm network (32x32 block per GPU), chi=8, d=2, Float64"


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm restated (oracle/bp_oracle.c), all host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_arm(p, seconds: float, max_steps: int = 1000, warmup: int = 1):
    """Times synchronous sweeps of the C oracle (absorption order, OpenMP over edges).  If one full sweep
    exceeds the budget, a random sample of edges is timed instead.  -> (updates/s, cores, sample, ms/step, steps)"""
    from oracle.c_oracle import COracle

    co = COracle(p.ga, p.phys_dim, p.link_dim, p.tensors, p.dtype)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would silently serialise the arm)
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    flat = co.pack(p.messages)
    ne = p.ga.ne
    # probe cost on a small sample
    rng = np.random.default_rng(0)
    probe = np.sort(rng.choice(ne, size=min(ne, 4 * cores), replace=False))
    co.sweep_jacobi(flat, edges=probe, nthreads=cores)  # cold (thread pool start-up, page faults)
    t0 = time.perf_counter()
    co.sweep_jacobi(flat, edges=probe, nthreads=cores)
    per_update = (time.perf_counter() - t0) / len(probe)
    full = per_update * ne
    if full <= seconds / 3:
        edges, n_upd, sample = None, ne, f"full sweeps of all {ne} directed edges"
    else:
        k = max(cores, int(seconds / 3 / per_update))
        edges = np.sort(rng.choice(ne, size=min(ne, k), replace=False))
        n_upd, sample = len(edges), f"{len(edges)} randomly sampled directed edges of {ne} per step"
    for _ in range(warmup):
        co.sweep_jacobi(flat, edges=edges, nthreads=cores)
    times = []
    t_end = time.perf_counter() + seconds
    while len(times) < max_steps and (time.perf_counter() < t_end or len(times) < 2):
        t0 = time.perf_counter()
        co.sweep_jacobi(flat, edges=edges, nthreads=cores)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return n_upd / (ms * 1e-3), cores, sample, ms, len(times)


def cpu_reference_arm_single(seconds: float, max_steps: int = 1000, dims=(64, 64)):
    """Single-layer workloads: the numpy oracle's synchronous sweep (one thread, per-edge contractions like the reference)
    on a `dims` sub-lattice of the workload.  -> (updates/s, cores, sample, ms/step, steps)"""
    from itnn_b200 import problems

    o = entry.import_oracle()
    q = problems.synthetic_ising(dims)
    tensors, msgs = problems.unpacked(q)
    op = o.make_problem(q.ga, tensors, "single")
    o.sweep_jacobi(op, msgs)
    times = []
    t_end = time.perf_counter() + seconds
    while len(times) < max_steps and (time.perf_counter() < t_end or len(times) < 2):
        t0 = time.perf_counter()
        msgs = o.sweep_jacobi(op, msgs)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    sample = f"{dims[0]}x{dims[1]} periodic sub-lattice of the workload ({q.ga.ne} directed edges per step), numpy oracle, one thread"
    return q.ga.ne / (ms * 1e-3), 1, sample, ms, len(times)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    entry.import_package()
    p, _, desc = build_workload(args.workload, 1)
    if p.mode == "single":
        budget = max(10.0, min(120.0, 4.0 * args.steps))
        val, cores, sample, ms, steps = cpu_reference_arm_single(budget, max_steps=args.steps)
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "schedule": "synchronous", "note": "reference CPU path restated in numpy (oracle/bp_oracle.py); "
                       "the Julia reference itself cannot run in this image"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    if p.tensors is None:  # cfg5: the host cannot stage 63 GiB; same buckets on an 8x8 sub-lattice
        from itnn_b200 import graphs, problems

        p = problems.make_config("cfg5", graph=graphs.named_grid((8, 8)))
        desc += " (CPU arm: 8x8 sub-lattice sample)"
    budget = max(10.0, min(120.0, 4.0 * args.steps))
    val, cores, sample, ms, steps = cpu_reference_arm(p, budget, max_steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if p.dtype.kind != "c" else "c128", "data": "synthetic",
        "config": {"workload": desc, "schedule": "synchronous", "note": "reference CPU path restated in C (oracle/bp_oracle.c, "
                   "absorption order, OpenMP over edges); the Julia reference itself cannot run in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, n: int = 4096, reps: int = 5) -> float:
    """cuBLAS DGEMM n^3 (torch.matmul, float64), best of `reps`, TFLOP/s -- the same method the driver used
    for the bf16 entry of MEASURED_PEAKS.json, which has no FP64 figure."""
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) * 1e-12


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    pkg = entry.import_package()
    from itnn_b200 import _lib, problems

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        try:  # keep the ranks' launch threads off each other's cores
            ncpu = os.cpu_count() or 1
            per = max(1, ncpu // world)
            if os.environ.get("BENCH_CORES_PER_RANK"):  # experiments: emulate the cores a rank gets at a larger N
                per = int(os.environ["BENCH_CORES_PER_RANK"])
            os.sched_setaffinity(0, set(range(local_rank * per, min(ncpu, (local_rank + 1) * per))))
        except Exception:
            pass
        # host-side plumbing only (IPC-handle exchange, barriers, max of the timings): gloo.  The data path
        # (cut-edge messages, residual) goes over NVLink peer memory inside libbpx, not through a collective library.
        dist.init_process_group("gloo")

    p, owner, desc = build_workload(args.workload, world)
    ctx = pkg.BPXContext(local_rank)
    # a dedicated non-default stream: torch events and the library's launches share it (handle 0, the
    # legacy default stream, would mean "library-internal stream" to bpx_set_stream)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_graph(p.ga.src, p.ga.dst, p.ga.slot, p.ga.nv)
    if args.kernel:
        ctx.set_kernel_policy(args.kernel)
    ctx.set_dims(p.dtype, p.mode, p.phys_dim if p.mode == "norm" else None, p.link_dim)
    if world > 1:  # partition first: site tensors are then allocated / generated for the owned vertices only
        from itnn_b200 import partition

        partition.connect(ctx, owner, rank, world)
    if p.tensors is None:  # inputs generated on the device (same recipe; the host cannot stage 63 GiB)
        ctx.fill_synthetic(123)
        flat0 = ctx.get_messages_flat()
    else:
        ctx.set_site_tensors(p.tensors)
        flat0 = ctx.pack_messages(p.messages)
        ctx.set_messages(flat0)
    n_local_updates = sum(b["edges"] for b in ctx.buckets())
    n_total_updates = p.ga.ne

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    # ---- value: everything resident, device-timed per step, L2 flushed between steps -------------------
    for _ in range(args.warmup):
        ctx.sweep_async(1)
    barrier()
    ctx.counters(reset=True)
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one NVML poller per box is enough (and several perturb the launch path of every GPU)
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        l2_flush.zero_()
        if args.flush == "write+read":
            l2_flush.view(torch.int64).sum()
        if world > 1:
            ctx.peer_barrier()  # untimed: align the ranks after their (untimed) L2 flushes, so that one rank's flush
            #                     does not sit inside its neighbour's timed sweep
        e0.record(stream)
        ctx.sweep_async(1)
        e1.record(stream)
    barrier()
    if args.dump_timing:
        import ctypes as C_
        buf = np.zeros(8 * 32 * 16, dtype=np.int64)
        ctx.lib.bpx_debug_timing.argtypes = [C_.c_void_p, C_.c_void_p, C_.c_int]
        ctx.lib.bpx_debug_timing(ctx.h, None, 0)  # allocate
        for _ in range(3):
            l2_flush.zero_()
            if world > 1:
                ctx.peer_barrier()
            ctx.sweep_async(1)
        barrier()
        ctx.lib.bpx_debug_timing(ctx.h, buf.ctypes.data_as(C_.c_void_p), buf.size)
        g = buf[2048:2048 + 8 * 148].reshape(148, 8)
        g = g[g[:, 0] > (1 << 50)]  # (the per-phase clock64 stamps of CTA 0 share the buffer)
        t0 = g[:, 0].min()
        sys.stderr.write(f"[rank {rank}] CTAs {len(g)}: start spread {g[:,0].max()-t0} ns; gate done at {np.median(g[:,1]-t0):.0f} (max {(g[:,1]-t0).max()}); "
                         f"compute end median {np.median(g[:,2]-t0):.0f} max {(g[:,2]-t0).max()}; epilogue end max {(g[:,3]-t0).max()}; "
                         f"after post: median {np.median(g[:,4]-t0):.0f} max {(g[:,4]-t0).max()} ns\n")
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    counters = ctx.counters()
    bucket_times = []
    for b, info in enumerate(ctx.buckets()):
        ms, n = ctx.bucket_time(b)
        bucket_times.append((info, ms, n))
    ctx.set_profiling(False)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_total_updates / (ms_per_step * 1e-3)
    residual = ctx.last_residual()

    # ---- roofline of the dominant bucket's kernel ----------------------------------------------------
    cplx = 4.0 if p.dtype.kind == "c" else 1.0
    w = p.dtype.itemsize
    all_buckets = ctx.buckets()
    dom_idx = max(range(len(bucket_times)), key=lambda i: bucket_times[i][1])
    dom, dom_ms, dom_n = bucket_times[dom_idx]
    z, chi, d = dom["degree"], dom["chi"], dom["phys"]
    # the dominant LAUNCH covers every bucket merged into it (bpx_bucket_info leader): sum their algorithmic work
    flops_per_launch = bytes_per_launch = 0.0
    merged = [b for b in all_buckets if b["leader"] == dom["leader"]]
    for b in merged:
        if p.mode == "single":  # vector messages; the factor shrinks by chi with every absorbed message
            flops_per_launch += b["edges"] * problems.single_layer_update_flops(b["degree"], b["chi"]) * cplx
            bytes_per_launch += b["vertices"] * float(b["chi"]) ** b["degree"] * w + 3.0 * b["edges"] * b["chi"] * w
            continue
        flops_per_launch += b["edges"] * 2.0 * b["degree"] * b["phys"] * float(b["chi"]) ** (b["degree"] + 1) * cplx
        bytes_per_launch += b["vertices"] * b["phys"] * float(b["chi"]) ** b["degree"] * w + 3.0 * b["edges"] * b["chi"] ** 2 * w
    avg_ms = dom_ms / max(dom_n, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp64_peak = measure_fp64_peak(torch)
    t_flop = flops_per_launch / (fp64_peak * 1e12)
    t_byte = bytes_per_launch / (hbm_peak * 1e9)
    if t_flop >= t_byte:
        roof = {"bound": "tensor", "achieved": flops_per_launch / (avg_ms * 1e-3) * 1e-12, "peak": fp64_peak, "unit": "TFLOP/s",
                "peak_source": "measured live: cuBLAS DGEMM 4096^3 via torch.matmul(float64), best of 5 (MEASURED_PEAKS.json has no FP64 entry)"}
    else:
        roof = {"bound": "hbm", "achieved": bytes_per_launch / (avg_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None
    if (args.workload, world) in PROFILED_TRAFFIC:
        roof["traffic"], roof["traffic_source"] = PROFILED_TRAFFIC[(args.workload, world)]
    elif dom["kernel"] == 3 and world == 1:
        roof["traffic"] = SLICED_TRAFFIC_PER_VERTEX * dom["vertices"]
        roof["traffic_source"] = "profiles/r1c_sliced_c16_ncu_summary.csv (per-vertex DRAM bytes x degree-4 vertices of this workload)"
    roof["kernel"] = {1: "bp_update_generic", 2: "bp_update_onchip", 3: "bp_update_sliced", 4: "bp_update_single_vertex"}.get(dom["kernel"], "?")
    roof["bucket"] = {"degree": z, "chi": chi, "phys": d, "updates_per_launch": sum(b["edges"] for b in merged),
                      "degrees_in_launch": sorted(b["degree"] for b in merged)}
    roof["avg_launch_ms"] = avg_ms
    roof["flops_per_launch"] = flops_per_launch
    roof["bytes_per_launch"] = bytes_per_launch
    roof["hbm_gbs_achieved"] = bytes_per_launch / (avg_ms * 1e-3) * 1e-9
    roof["share_of_step"] = dom_ms / max(sum(t[1] for t in bucket_times), 1e-30)

    # ---- e2e: host buffers through the C ABI, H2D + sweep + D2H inside the timed region -----------------
    e2e = None
    if not args.no_e2e:
        nbytes = flat0.nbytes
        pin_in = torch.empty(flat0.size, dtype=torch.float64 if p.dtype.kind != "c" else torch.complex128).pin_memory()
        pin_out = torch.empty_like(pin_in).pin_memory()
        h_in, h_out = pin_in.numpy(), pin_out.numpy()
        h_in[:] = flat0
        def e2e_step(src, dst):
            # one C-ABI call: upload this rank's messages from pinned memory, one sweep (cut-edge messages travel between
            # the devices), this rank's new messages + the (global) residual back in host memory
            return ctx.sweep_host(src, dst)

        for _ in range(2):
            e2e_step(h_in, h_out)
        barrier()
        ee = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            res_e2e = e2e_step(h_in, h_out)
            e1.record(stream)
            ee.append((e0, e1))
            h_in, h_out = h_out, h_in     # next step consumes this step's result
        barrier()
        e_ms = float(sum(a.elapsed_time(b) for a, b in ee))
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": n_total_updates / (e_ms / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(nbytes + 8), "ms_per_step": e_ms / args.steps,
               "what": "bpx_sweep_host, one C-ABI call per step with pinned host buffers: messages H2D + one sweep + messages and residual back in host "
                       "memory; site tensors resident.  Single-launch sweeps stream: the kernel runs while the upload arrives in chunks and "
                       "stores new messages straight into the host buffer (one CUDA-graph launch per step)"}

    # ---- CPU baseline (rank 0, N = 1) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pc, note = p, ""
        if p.mode == "single":
            v, cores, sample, ms, steps = cpu_reference_arm_single(args.cpu_seconds)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample}; {steps} steps of {ms:.1f} ms"}
        else:
            if p.tensors is None:  # cfg5: time the CPU on an 8x8 sub-lattice of the same buckets
                from itnn_b200 import graphs
                pc, note = problems.make_config("cfg5", graph=graphs.named_grid((8, 8))), "8x8 sub-lattice of the workload; "
            v, cores, sample, ms, steps = cpu_reference_arm(pc, args.cpu_seconds)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{note}{sample}; {steps} steps of {ms:.1f} ms"}

    # ---- optional: BP to convergence (the north star's end-to-end statement) --------------------------------
    conv = None
    if args.converge > 0.0:
        if p.tensors is None:
            ctx.fill_synthetic(123)  # back to the initial messages (device-side recipe)
        else:
            ctx.set_messages(flat0)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        c0.record(stream)
        res_c, done_c = ctx.sweep(200, args.converge)
        c1.record(stream)
        barrier()
        tw = time.perf_counter() - tw
        c_ms = c0.elapsed_time(c1)
        if world > 1:
            t = torch.tensor([c_ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c_ms = float(t.item())
        conv = {"tol": args.converge, "maxiter": 200, "sweeps": int(done_c), "residual": res_c, "ms": c_ms, "wall_s": tw,
                "updates_per_s": n_total_updates * int(done_c) / (c_ms * 1e-3),
                "what": "bpx_sweep(200, tol): synchronous sweeps until the fused (global) residual < tol, checked on the host after every sweep; L2 not flushed"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if "STRONG" in desc else "weak", "vs_baseline": None,
            "dtype": "f64" if p.dtype.kind != "c" else "c128", "data": "synthetic",
            "config": {"workload": desc, "schedule": "synchronous (Jacobi) sweep, sum-normalised, residual fused",
                       "updates_per_step": n_total_updates, "updates_per_gpu": n_local_updates,
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "timing": "CUDA events per step on the launching stream" + ("; ranks aligned by a device-side barrier after each flush" if world > 1 else "")},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(counters["launches"]), "residual_after_bench": residual,
            "wall_s_timed_region": t_wall, "convergence": conv,
            "buckets": [{"degree": i["degree"], "chi": i["chi"], "edges": i["edges"], "kernel": i["kernel"],
                         "ms_per_launch": ms / max(n, 1)} for i, ms, n in bucket_times],
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
