#!/usr/bin/env python
"""Benchmark of the BP message-update hot path (BASELINE.json metric: BP message updates/s).

  python bench.py --gpus N --steps K --warmup W [--workload cfg5] [--impl reference]

A "step" is ONE synchronous BP sweep (every directed edge updated once) over a synthetic PEPS norm network.

Default workload at EVERY N: BASELINE config 5, the north-star target -- the 256x256 square-lattice PEPS norm network
with chi = 16, d = 2, Float64 (261 120 updates per sweep, 63 GiB of site tensors; it fits one B200).  For N > 1 the
lattice is vertex-sharded into N strips of 256/N rows, one per GPU (STRONG scaling): cut-edge messages are pushed into
the neighbour ranks over NVLink peer memory inside the sweep kernels and the residual is max-reduced over the ranks.

One JSON line on rank 0 (see the task contract):
  value          updates/s with everything resident in HBM (CUDA events per step, L2 flushed between steps)
  e2e            the same through the C ABI with HOST buffers (message upload + sweep + message download per step)
  roofline       dominant kernel against the measured FP64 (DGEMM) / HBM peaks
  cpu_baseline   the CPU restatement of the reference algorithm (oracle/bp_oracle.c, all host cores) on a bounded sample
  parity         after the timed region: a random sample of directed edges (incl. cut edges and both sides of every rank
                 boundary) recomputed by the CPU oracle from the DEVICE's own inputs; the run FAILS above 1e-10
  convergence    BP to convergence (maxiter 200, StopWhenConverged(1e-10)) from the initial messages
  other_configs  the other BASELINE configs measured in the same process (cfg4 sharded over the same N; cfg1-3 at N = 1)

The oracle is used as the checker (parity) and as the CPU baseline only -- never on the measured path.
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
# sliced chi=16 kernel: DRAM bytes per degree-4 vertex and the capture they come from (profiles/sliced_traffic.json is
# rewritten whenever the kernel is re-profiled; fallback = round 1's capture)
SLICED_TRAFFIC_PER_VERTEX = ((926.93e6 + 415.05e6) / 196.0, "profiles/r1c_sliced_c16_ncu_summary.csv")
try:
    _st = json.load(open(os.path.join(ROOT, "profiles", "sliced_traffic.json")))
    SLICED_TRAFFIC_PER_VERTEX = (float(_st["dram_bytes_per_vertex"]), _st["source"])
except Exception:
    pass

METRIC = "bp_message_updates_per_s"
UNIT = "updates/s"
PARITY_TOL = 1e-10

DESC = {
    "cfg1": "4x4 square-lattice PEPS norm network, chi=2, d=2, Float64",
    "cfg2": "32x32 square-lattice PEPS norm network, chi=8, d=2, Float64",
    "cfg3": "heavy-hex 127-site PEPS norm network, chi=16, d=2, ComplexF64",
    "cfg4": "16x16x16 periodic cubic PEPS norm network, chi=4, d=2, Float64",
    "cfg5": "256x256 square-lattice PEPS norm network, chi=16, d=2, Float64",
    "cfg2c": "32x32 square-lattice PEPS norm network, chi=8, d=2, ComplexF64 (complex twin of config 2; not a BASELINE config)",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg1", "cfg2", "cfg2w", "cfg3", "cfg4", "cfg5", "cfg5s", "cfg2c", "ising", "apply"])
    ap.add_argument("--kernel", type=int, default=0, help="force a kernel family (include/bpx.h BPX_KERNEL_*)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (other_configs)")
    ap.add_argument("--no-beliefs", action="store_true")
    ap.add_argument("--converge", type=float, default=1e-10, metavar="TOL",
                    help="BP to convergence (maxiter 200, StopWhenConverged(TOL)) from the initial messages, reported under "
                         "\"convergence\" (untimed by the step metric); 0 disables")
    ap.add_argument("--parity-edges", type=int, default=64, help="random directed edges per rank recomputed by the oracle (+ boundary edges)")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all N GPUs from ONE process through bpx_create_multi (no torchrun); the default for N > 1 is one "
                         "process per GPU, the launch mode the driver uses")
    ap.add_argument("--dump-timing", action="store_true", help="debug (BPX_ONCHIP_TIMING builds): per-CTA globaltimer stamps of the last sweep")
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between timed steps: 256 MiB memset, optionally followed by a read pass over the same buffer "
                         "(leaves L2 full of clean instead of dirty lines)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def build_workload(name: str, world: int, host_data=None):
    """-> (SyntheticProblem, owner list or None, description, scaling)."""
    from itnn_b200 import graphs, problems

    if name == "ising":  # not a BASELINE config: the HBM-bound single-layer bucket of the path
        if world > 1:
            raise SystemExit("--workload ising is a single-GPU workload")
        return (problems.make_config("ising"), None,
                "1024x1024 periodic square-lattice Ising partition-function network (ising_network recipe, beta=0.3), single layer, chi=2, Float64",
                "weak")
    if name == "cfg5s":  # cfg5's buckets on a lattice that is quick to generate (profiling runs)
        g = graphs.named_grid((24, 24))
        p = problems.make_config("cfg5", graph=g)
        return p, None, "24x24 square-lattice PEPS norm network (cfg5 buckets), chi=16, d=2, Float64", "weak"
    if name == "cfg2w":  # round 1's default: weak scaling of cfg2 blocks (not a BASELINE config for N > 1)
        g = graphs.named_grid((32, 32 * world))
        p = problems.make_config("cfg2", graph=g)
        owner = [(v[1] - 1) // 32 for v in p.ga.vertices] if world > 1 else None
        return p, owner, f"32x{32 * world} square-lattice PEPS norm network (32x32 block per GPU), chi=8, d=2, Float64", "weak"
    if host_data is None:
        host_data = name != "cfg5"
    p = problems.make_config(name, host_data=host_data)
    owner = None
    if world > 1:
        if name == "cfg5":  # the north-star target: 256x256 vertex-sharded into strips of rows
            owner = [(v[1] - 1) * world // 256 for v in p.ga.vertices]
        elif name == "cfg4":  # BASELINE config 4 "at 1/2/4/8 B200": slabs of the periodic cube
            owner = [(v[2] - 1) * world // 16 for v in p.ga.vertices]
        elif name in ("cfg2", "cfg2c"):
            owner = [(v[1] - 1) * world // 32 for v in p.ga.vertices]
        else:
            raise SystemExit(f"--workload {name} is a single-GPU workload (replicas only)")
    return p, owner, DESC[name], "strong"


def shard_note(name: str, world: int) -> str:
    if world == 1:
        return "one GPU"
    return {"cfg5": f"{world} strips of {256 // world} rows", "cfg4": f"{world} slabs of {16 // world} planes",
            "cfg2": f"{world} strips of {32 // world} rows", "cfg2c": f"{world} strips of {32 // world} rows",
            "cfg2w": f"one 32x32 block per GPU ({world} blocks)"}.get(name, f"{world} blocks")


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
def cpu_reference_arm(p, seconds: float, max_steps: int = 1000, warmup: int = 1, edges_fixed=None):
    """Times synchronous sweeps of the C oracle (absorption order, OpenMP over edges).  If one full sweep
    exceeds the budget, a random sample of edges is timed instead; `edges_fixed` = (edge list, description) times exactly
    those edges per step.  -> (updates/s, cores, sample, ms/step, steps)"""
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
    if edges_fixed is not None:
        edges, sample = np.asarray(edges_fixed[0], dtype=np.int64), edges_fixed[1]
        n_upd = len(edges)
    elif full <= seconds / 3:
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


def cpu_problem(name: str, p):
    """The problem the CPU arms time: `p` itself, or -- for workloads whose inputs the host cannot stage (cfg5: 63 GiB) --
    edges of an 8x8 sub-lattice drawn in the workload's own degree mix (cfg5: 98.8 % of the updates leave degree-4
    vertices, 16x the work of a degree-3 update; a plain 8x8 sweep would be 36 % boundary updates and flatter the CPU).
    -> (problem, edges_fixed or None)"""
    if p.tensors is not None:
        return p, None
    from itnn_b200 import graphs, problems

    q = problems.make_config("cfg5", graph=graphs.named_grid((8, 8)))
    deg_full = np.diff(np.asarray(p.ga.row_ptr))
    deg_q = np.diff(np.asarray(q.ga.row_ptr))
    src_q = np.asarray(q.ga.src)
    share = {int(z): float((deg_full == z).sum() * z) / p.ga.ne for z in np.unique(deg_full)}  # share of the updates per source degree
    z_top = max(share, key=share.get)
    pool_top = np.nonzero(deg_q[src_q] == z_top)[0]
    edges = list(pool_top)
    mix = {z_top: len(pool_top)}
    for z, sh in share.items():
        if z == z_top:
            continue
        k = int(round(len(pool_top) * sh / share[z_top]))
        pool = np.nonzero(deg_q[src_q] == z)[0]
        edges += list(pool[:k])
        mix[z] = min(k, len(pool))
    note = ("edges of an 8x8 sub-lattice (same chi=16 buckets) in the workload's degree mix: " +
            ", ".join(f"{n} updates out of degree-{z} vertices" for z, n in sorted(mix.items(), reverse=True)) + " per step")
    return q, (np.sort(np.asarray(edges, dtype=np.int64)), note)


def run_reference(args, rank: int, world: int):
    """Reference arm: the reference's CPU algorithm (restated in C, oracle/bp_oracle.c -- the Julia package cannot run in
    this image) on all host cores, on a bounded sample of the SAME workload.  The product library is not loaded: lattice
    builders are pure Python and the synthetic inputs come from the numpy restatement of the shared RNG."""
    if rank != 0:
        return
    entry.import_package()  # pure-Python modules only; libbpx.so is loaded lazily and nothing below triggers it
    from itnn_b200 import problems
    from oracle import synthetic_rng

    problems.fill_randn = synthetic_rng.fill_randn
    p, _, desc, scaling = build_workload(args.workload, 1 if args.workload != "cfg2w" else world, host_data=None)
    budget = max(10.0, min(120.0, 4.0 * args.steps))
    if p.mode == "single":
        val, cores, sample, ms, steps = cpu_reference_arm_single(budget, max_steps=args.steps)
        dtype, note = "f64", "reference CPU path restated in numpy (oracle/bp_oracle.py)"
    else:
        pc, fixed = cpu_problem(args.workload, p)
        val, cores, sample, ms, steps = cpu_reference_arm(pc, budget, max_steps=args.steps, warmup=max(1, min(args.warmup, 3)), edges_fixed=fixed)
        dtype = "f64" if p.dtype.kind != "c" else "c128"
        note = "reference CPU path restated in C (oracle/bp_oracle.c, absorption order, OpenMP over edges)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": desc, "schedule": "synchronous (Jacobi) sweep, sum-normalised, residual fused"},
        "note": note + "; the Julia reference itself cannot run in this image; each step is the bounded sample named in cpu_baseline.sample",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "libbpx_loaded": bool(sys.modules["itnn_b200"]._lib._lib is not None),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------------
_FP64_PEAKS = None


def measure_fp64_peaks(torch, n: int = 8192, reps: int = 5, sustained_s: float = 2.0):
    """cuBLAS DGEMM n^3 (torch.matmul, float64): best of `reps` (burst) and back to back for `sustained_s` seconds
    (sustained) -- the method the driver used for the bf16 entries of MEASURED_PEAKS.json, which has no FP64 figure.
    A committed copy of one such measurement on this pool is profiles/fp64_peak.json (tools/fp64_peak.py)."""
    global _FP64_PEAKS
    if _FP64_PEAKS is not None:
        return _FP64_PEAKS
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    k = max(4, int(sustained_s / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    del a, b, c
    torch.cuda.empty_cache()
    _FP64_PEAKS = {"burst": flop / (best * 1e-3) * 1e-12, "sustained": flop * k / (e0.elapsed_time(e1) * 1e-3) * 1e-12,
                   "how": f"measured live: cuBLAS DGEMM {n}^3 via torch.matmul(float64), best of {reps} (burst) / {k} back to back (sustained); "
                          "MEASURED_PEAKS.json has no FP64 entry (committed copy of the same measurement: profiles/fp64_peak.json)"}
    return _FP64_PEAKS


class Dist:
    """The host-side plumbing of a multi-process run (gloo): IPC-handle exchange, barriers, reductions of timings."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def reduce(self, x: float, op: str = "max") -> float:
        if not self.dist:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def gather_objects(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def parity_check(ctx, p, owner, D: Dist, n_random: int, seed: int = 7):
    """Recompute a sample of directed edges with the CPU oracle from the DEVICE's own inputs (site tensor of the source
    vertex, previous-sweep messages as they sit in this rank's memory, halo messages included) and compare with what one
    more sweep produced (beliefpropagation.jl:242-257, messagecache.jl:128-131).  Multi-GPU: every rank samples its own
    edges -- random ones, edges ACROSS each of its cuts, and edges out of the boundary vertices on its side that consume
    a message received from the peer -- and the delivered copies of cut-edge messages are compared bit for bit with the
    sender's.  -> dict (max over ranks)."""
    o = entry.import_oracle()
    ga = p.ga
    rank, world = D.rank, D.world
    src = np.asarray(ga.src)
    dst = np.asarray(ga.dst)
    rev = np.asarray(ga.rev)
    row_ptr = np.asarray(ga.row_ptr)
    own = np.zeros(ga.nv, dtype=np.int64) if owner is None else np.asarray(owner, dtype=np.int64)
    before = ctx.get_messages_flat()
    res, _ = ctx.sweep(1)
    after = ctx.get_messages_flat()
    rng = np.random.default_rng(seed + rank)
    mine = np.nonzero(own[src] == rank)[0]
    sample = set(rng.choice(mine, size=min(n_random, len(mine)), replace=False).tolist()) if len(mine) else set()
    n_cut = n_halo = 0
    if world > 1:
        cut_out = mine[own[dst[mine]] != rank]                      # owned edges whose head lives on a peer
        recv_v = np.unique(dst[(own[dst] == rank) & (own[src] != rank)])  # my vertices that receive a peer's message
        for peer in np.unique(own[dst[cut_out]]):
            c = cut_out[own[dst[cut_out]] == peer]
            pick = rng.choice(c, size=min(8, len(c)), replace=False)
            sample.update(pick.tolist())
            n_cut += len(pick)
        if len(recv_v):
            for v in rng.choice(recv_v, size=min(12, len(recv_v)), replace=False):
                # an out-edge of v that stays on my side consumes the halo message
                outs = [e for e in range(row_ptr[v], row_ptr[v + 1]) if own[dst[e]] == rank]
                if outs:
                    sample.add(int(outs[0]))
                    n_halo += 1
    mo = ctx.msg_off
    ld = ctx.link_dim

    def msg(flat, e):
        chi = int(ld[e])
        return flat[mo[e]:mo[e + 1]].reshape((chi, chi) if p.mode == "norm" else (chi,), order="F")

    worst = 0.0
    worst_res = 0.0
    for e in sorted(sample):
        u = int(src[e])
        z = int(row_ptr[u + 1] - row_ptr[u])
        dims = [int(ld[f]) for f in range(row_ptr[u], row_ptr[u + 1])]
        A = ctx.get_site_tensor(u)
        ins = [None if f == e else msg(before, int(rev[f])) for f in range(row_ptr[u], row_ptr[u + 1])]
        slot = int(e - row_ptr[u])
        if p.mode == "norm":
            A = A.reshape([int(p.phys_dim[u])] + dims, order="F")
            want = o.normalize_message(o.contract_norm(A, slot, ins))
        else:
            A = A.reshape(dims, order="F")
            want = o.normalize_message(o.contract_single(A, slot, ins))
        got = msg(after, e)
        worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
        worst_res = max(worst_res, o.edge_residual(msg(before, e), want))
    # the fused residual is the maximum over ALL edges: it can never be below the oracle's value on the sample
    res_ok = bool(res + 1e-12 >= worst_res)
    halo_mismatch = 0
    halo_checked = 0
    if world > 1:
        sent = {}
        for peer in np.unique(own[dst[cut_out]]):
            c = cut_out[own[dst[cut_out]] == peer]
            for e in rng.choice(c, size=min(32, len(c)), replace=False):
                sent[int(e)] = after[mo[e]:mo[e + 1]].tobytes()
        for q, d in enumerate(D.gather_objects(sent)):
            if q == rank:
                continue
            for e, raw in d.items():
                if own[dst[e]] == rank:
                    halo_checked += 1
                    if after[mo[e]:mo[e + 1]].tobytes() != raw:
                        halo_mismatch += 1
    out = {"max_rel_err": D.reduce(worst, "max"), "n_edges": int(D.reduce(float(len(sample)), "sum")), "tol": PARITY_TOL,
           "cut_edges_checked": int(D.reduce(float(n_cut), "sum")), "halo_consumers_checked": int(D.reduce(float(n_halo), "sum")),
           "halo_copies_compared": int(D.reduce(float(halo_checked), "sum")), "halo_copies_mismatched": int(D.reduce(float(halo_mismatch), "sum")),
           "residual_consistent": bool(D.reduce(0.0 if res_ok else 1.0, "max") == 0.0), "residual_of_checked_sweep": res,
           "what": "after the timed region: one more sweep; sampled directed edges recomputed by the CPU oracle (numpy, absorption order) from the "
                   "device's own site tensors and previous-sweep messages; max |got - want| / max |want| over the sample"}
    out["ok"] = bool(out["max_rel_err"] < PARITY_TOL and out["halo_copies_mismatched"] == 0 and out["residual_consistent"])
    return out


def measure(args, name, pkg, torch, D: Dist, local_rank, stream, l2_flush, main: bool):
    """One workload on this process group: resident value (+ roofline), e2e, parity, convergence.  -> dict (rank 0's view)"""
    from itnn_b200 import problems

    rank, world = D.rank, D.world
    p, owner, desc, scaling = build_workload(name, world)
    ctx = pkg.BPXContext(local_rank)
    # a dedicated non-default stream: torch events and the library's launches share it (handle 0, the
    # legacy default stream, would mean "library-internal stream" to bpx_set_stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_graph(p.ga.src, p.ga.dst, p.ga.slot, p.ga.nv)
    if args.kernel and main:
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
    steps, warmup = args.steps, args.warmup

    def barrier():
        D.barrier()
        torch.cuda.synchronize()

    # ---- value: everything resident, device-timed per step, L2 flushed between steps -------------------
    for _ in range(warmup):
        ctx.sweep_async(1)
    barrier()
    ctx.counters(reset=True)
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    if rank == 0 and main:  # one NVML poller per box is enough (and several perturb the launch path of every GPU)
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
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
    if args.dump_timing and main:
        dump_timing(ctx, torch, l2_flush, world, rank, barrier)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 and main else None
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = D.reduce(float(step_ms.sum()), "max")
    counters = ctx.counters()
    bucket_times = []
    for b, info in enumerate(ctx.buckets()):
        ms, n = ctx.bucket_time(b)
        bucket_times.append((info, ms, n))
    ctx.set_profiling(False)
    ms_per_step = total_ms / steps
    value = n_total_updates / (ms_per_step * 1e-3)
    residual = ctx.last_residual()

    # ---- roofline of the dominant bucket's kernel (this rank's launch; rank 0 reports) -------------------
    cplx = 4.0 if p.dtype.kind == "c" else 1.0
    w = p.dtype.itemsize
    all_buckets = ctx.buckets()
    dom_idx = max(range(len(bucket_times)), key=lambda i: bucket_times[i][1])
    dom, dom_ms, dom_n = bucket_times[dom_idx]
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
    fp = measure_fp64_peaks(torch)
    # a launch of tens of milliseconds inside a seconds-long loop runs under the power cap: sustained figure; short launches
    # timed one at a time: burst figure (B200_PROFILING.md)
    long_kernel = avg_ms >= 10.0
    fp64_peak = fp["sustained"] if long_kernel else fp["burst"]
    t_flop = flops_per_launch / (fp64_peak * 1e12)
    t_byte = bytes_per_launch / (hbm_peak * 1e9)
    if t_flop >= t_byte:
        roof = {"bound": "tensor", "achieved": flops_per_launch / (avg_ms * 1e-3) * 1e-12, "peak": fp64_peak, "unit": "TFLOP/s",
                "peak_source": ("sustained, " if long_kernel else "burst, ") + fp["how"], "fp64_burst": fp["burst"], "fp64_sustained": fp["sustained"]}
    else:
        roof = {"bound": "hbm", "achieved": bytes_per_launch / (avg_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None
    if (name, world) in PROFILED_TRAFFIC:
        roof["traffic"], roof["traffic_source"] = PROFILED_TRAFFIC[(name, world)]
    elif dom["kernel"] == 3:
        roof["traffic"] = SLICED_TRAFFIC_PER_VERTEX[0] * dom["vertices"]
        roof["traffic_source"] = SLICED_TRAFFIC_PER_VERTEX[1] + " (per-vertex DRAM bytes x degree-4 vertices of this launch)"
    roof["kernel"] = {1: "bp_update_generic", 2: "bp_update_onchip", 3: "bp_update_sliced", 4: "bp_update_single_vertex"}.get(dom["kernel"], "?")
    roof["bucket"] = {"degree": dom["degree"], "chi": dom["chi"], "phys": dom["phys"], "updates_per_launch": sum(b["edges"] for b in merged),
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
        for _ in range(2):
            ctx.sweep_host(h_in, h_out)
        barrier()
        ee = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            # one C-ABI call: upload this rank's messages from pinned memory, one sweep (cut-edge messages travel between
            # the devices), this rank's new messages + the (global) residual back in host memory
            ctx.sweep_host(h_in, h_out)
            e1.record(stream)
            ee.append((e0, e1))
            h_in, h_out = h_out, h_in     # next step consumes this step's result
        barrier()
        e_ms = D.reduce(float(sum(a.elapsed_time(b) for a, b in ee)), "max")
        own_bytes = int(nbytes * n_local_updates / max(n_total_updates, 1)) if world > 1 else int(nbytes)
        e2e = {"value": n_total_updates / (e_ms / steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": own_bytes,
               "d2h_bytes_per_step": own_bytes + 8, "ms_per_step": e_ms / steps,
               "what": "bpx_sweep_host, one C-ABI call per step with pinned host buffers: this rank's messages H2D + one sweep + its new messages "
                       "and the (global) residual back in host memory; site tensors resident (bytes are per rank).  Sweeps made of specialised "
                       "launches stream: the kernel runs while the upload arrives in chunks and stores new messages straight into the host buffer"}
        del pin_in, pin_out, h_in, h_out

    # ---- parity: sampled edges against the CPU oracle, from the device's own inputs ------------------------
    parity = None
    if not args.no_parity:
        parity = parity_check(ctx, p, owner, D, args.parity_edges)

    # ---- BP to convergence (the north star's end-to-end statement) -----------------------------------------
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
        c_ms = D.reduce(c0.elapsed_time(c1), "max")
        conv = {"tol": args.converge, "maxiter": 200, "sweeps": int(done_c), "residual": res_c, "ms": c_ms, "wall_s": tw,
                "updates_per_s": n_total_updates * int(done_c) / (c_ms * 1e-3), "converged": bool(res_c < args.converge),
                "what": "bpx_sweep(200, tol) from the initial messages: synchronous sweeps until the fused (global) residual < tol "
                        "(beliefpropagation.jl:46-54, 69-92); L2 not flushed"}

    # ---- beliefs: bethe_free_energy's device part (messagecache.jl:139-201) ---------------------------------
    beliefs = None
    if main and not args.no_beliefs and world == 1:
        try:
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            f = ctx.bethe_free_energy()
            b1.record(stream)
            barrier()
            beliefs = {"ms": b0.elapsed_time(b1), "sweep_times": b0.elapsed_time(b1) / ms_per_step, "log_z_bp": float(np.real(f)),
                       "promoted_to_complex": isinstance(f, complex),
                       "what": "bpx_bethe_free_energy (messagecache.jl:185-201): vertex scalars on the update kernels, edge scalars and the "
                               "log-sum reduction on the device; 7 doubles come back"}
        except Exception as ex:  # noqa: BLE001
            beliefs = {"error": str(ex)}

    # ---- CPU baseline (rank 0, N = 1) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        secs = args.cpu_seconds if main else min(args.cpu_seconds, 3.0)
        if p.mode == "single":
            v, cores, sample, ms, st = cpu_reference_arm_single(secs)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample}; {st} steps of {ms:.1f} ms"}
        else:
            pc, fixed = cpu_problem(name, p)
            v, cores, sample, ms, st = cpu_reference_arm(pc, secs, edges_fixed=fixed)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample}; {st} steps of {ms:.1f} ms"}

    out = {
        "value": value, "ms_per_step": ms_per_step, "scaling": scaling, "dtype": "f64" if p.dtype.kind != "c" else "c128",
        "config": {"workload": desc, "sharding": shard_note(name, world), "schedule": "synchronous (Jacobi) sweep, sum-normalised, residual fused",
                   "updates_per_step": n_total_updates, "updates_per_gpu": n_local_updates,
                   "l2": "flushed between timed steps (256 MiB memset)",
                   "timing": "CUDA events per step on the launching stream" + ("; ranks aligned by a device-side barrier after each flush; max over ranks" if world > 1 else "")},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "parity": parity, "convergence": conv, "beliefs": beliefs,
        "gpu_launches": int(counters["launches"]), "residual_after_bench": residual, "wall_s_timed_region": t_wall,
        "buckets": [{"degree": i["degree"], "chi": i["chi"], "edges": i["edges"], "kernel": i["kernel"],
                     "ms_per_launch": ms / max(n, 1)} for i, ms, n in bucket_times],
    }
    ctx.close()
    return out


def dump_timing(ctx, torch, l2_flush, world, rank, barrier):
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


def brief(r):
    """What an `other_configs` entry keeps of a measurement."""
    keep = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": r["scaling"], "dtype": r["dtype"],
            "workload": r["config"]["workload"], "sharding": r["config"]["sharding"], "updates_per_step": r["config"]["updates_per_step"],
            "roofline": {k: r["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "avg_launch_ms", "traffic")},
            "e2e": None if r["e2e"] is None else {k: r["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "parity": None if r["parity"] is None else {k: r["parity"][k] for k in ("max_rel_err", "n_edges", "tol", "halo_copies_compared", "halo_copies_mismatched", "ok")},
            "convergence": None if r["convergence"] is None else {k: r["convergence"][k] for k in ("sweeps", "residual", "ms", "converged")},
            "cpu_baseline": None if r["cpu_baseline"] is None else {k: r["cpu_baseline"][k] for k in ("value", "cores", "kind")},
            "gpu_launches": r["gpu_launches"]}
    return keep


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch

    pkg = entry.import_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    D = Dist(rank, world)
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
        D.dist.init_process_group("gloo")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    r = measure(args, args.workload, pkg, torch, D, local_rank, stream, l2_flush, main=True)
    others = {}
    if not args.no_others and args.workload == "cfg5":
        names = ["cfg4"] if world > 1 else ["cfg1", "cfg2", "cfg3", "cfg4"]
        if world > 1 and 16 % world:
            names = []
        for nm in names:
            try:
                others[nm] = brief(measure(args, nm, pkg, torch, D, local_rank, stream, l2_flush, main=False))
            except Exception as ex:  # noqa: BLE001 -- never lose the headline line to a side measurement
                others[nm] = {"error": f"{type(ex).__name__}: {ex}"}
    ok = True
    if rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"], "vs_baseline": None,
            "dtype": r["dtype"], "data": "synthetic", "config": r["config"],
            "roofline": r["roofline"], "cpu_baseline": r["cpu_baseline"], "e2e": r["e2e"], "clocks": r["clocks"],
            "gpu_launches": r["gpu_launches"], "parity": r["parity"], "convergence": r["convergence"], "beliefs": r["beliefs"],
            "residual_after_bench": r["residual_after_bench"], "wall_s_timed_region": r["wall_s_timed_region"],
            "buckets": r["buckets"], "other_configs": others,
        }
        print(json.dumps(line), flush=True)
        for nm, res in [(args.workload, r)] + list(others.items()):
            par = res.get("parity") if isinstance(res, dict) else None
            if par is not None and not par.get("ok", True):
                sys.stderr.write(f"bench.py: PARITY FAILURE on {nm}: {json.dumps(par)}\n")
                ok = False
            # the synthetic BASELINE workloads converge in tens of sweeps: a run that hits maxiter has a handful of wrong
            # messages somewhere that a 64-edge sample can miss (this caught a protocol race in round 2)
            cv = res.get("convergence") if isinstance(res, dict) else None
            if cv is not None and not cv.get("converged", True):
                sys.stderr.write(f"bench.py: BP DID NOT CONVERGE on {nm} within maxiter: {json.dumps(cv)}\n")
                ok = False
    if world > 1:
        D.dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def run_apply(args):
    """`--workload apply`: layers of vertex-disjoint two-site gates (the four matchings of
    a 64x64 chi = 16 PEPS = one Trotter step of a nearest-neighbour Hamiltonian) through bpx_apply_two_site_gates.  A step =
    one layer.  Not BASELINE.json's metric (that is the default run); one JSON line with the gate path's own metric."""
    import torch

    pkg = entry.import_package()
    from itnn_b200 import graphs, problems

    nx = ny = 64
    chi, d = 16, 2
    p = problems.synthetic_peps(graphs.named_grid((nx, ny)), chi, d, np.float64)
    ga = p.ga
    rng = np.random.default_rng(0)
    vid = {v: i for i, v in enumerate(ga.vertices)}
    layers = []
    for axis in (0, 1):
        for parity in (0, 1):
            es = []
            for x in range(1, nx + 1):
                for y in range(1, ny + 1):
                    w = (x + 1, y) if axis == 0 else (x, y + 1)
                    if (x if axis == 0 else y) % 2 == parity and w in vid:
                        es.append(ga.edge_index[(vid[(x, y)], vid[w])])
            layers.append(es)
    dd = d * d

    def ops_for(n):
        o = rng.standard_normal((n, dd * dd))
        return [(np.eye(dd).ravel() + 0.1 * o[i]).reshape((d,) * 4, order="F") for i in range(n)]

    stream = torch.cuda.current_stream()
    mon = ClockSampler(0)
    with pkg.BPXContext(0) as ctx:
        problems.upload(ctx, p)
        ctx.sweep(3, 0.0, True)
        times, gates = [], 0
        for it in range(args.warmup + args.steps):
            es = layers[it % 4]
            ops = ops_for(len(es))
            if it == args.warmup:
                mon.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record(stream)
            ctx.apply_two_site_gates(es, ops, max_rank=chi, normalize=True)  # synchronous C-ABI call on the library's stream
            e1.record(stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
                gates += len(es)
        clocks = mon.stop()
        res, _ = ctx.sweep(1, 0.0, True)
        stats = ctx.apply_stats()
        launches = ctx.counters()["launches"]
    value = gates / sum(times)
    # dense work of one bulk gate on the Gram path: per tensor absorb 3 rows cols chi + Gram + final rows cols^2
    rows, cols = chi ** 3, d * chi
    flops_per_gate = 2.0 * 2.0 * (3 * rows * cols * chi + 2 * rows * cols * cols)
    peaks = measure_fp64_peaks(torch)
    out = {
        "metric": "bp_simple_update_gates_per_s", "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{nx}x{ny} square-lattice PEPS, chi={chi}, d={d}, Float64: gate layers = the four matchings (one Trotter step)",
                   "gates_per_step": [len(l) for l in layers],
                   "timing": "host wall clock around the synchronous C-ABI call (operator upload, descriptors, kernel, status read-back)"},
        "roofline": {"bound": "tensor", "achieved": value * flops_per_gate * 1e-12, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": value * flops_per_gate * 1e-12 / peaks["sustained"], "traffic": (6.28e9 / 592 + 2.47e9 / 592) * sum(len(l) for l in layers) / 4,
                     "traffic_source": "DRAM bytes of one launch over 592 gates: side kernel 6.28 GB (profiles/r2bq_apply3_sides_ncu_summary.csv)"
                                       " + final kernel 2.47 GB (r2av_apply3_final_ncu_summary.csv; same traffic after its rewrite) + bond "
                                       "kernel 0.1 GB, per gate x gates of a layer",
                     "kernel": "bp_apply3_sides + bp_apply3_bond + bp_apply3_final", "flops_per_gate": flops_per_gate, "peak_source": peaks["how"]},
        "cpu_baseline": None,
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": int(np.mean([len(l) for l in layers]) * dd * dd * 8),
                "d2h_bytes_per_step": int(np.mean([len(l) for l in layers]) * (chi * 8 + 4)),
                "what": "the same call: gate operators from host memory in, singular values and per-gate status back; state and "
                        "environment stay resident (they are the iterate of a simple-update evolution)"},
        "clocks": clocks, "gpu_launches": int(launches),
        "gates_on_gram_kernel": stats[0], "gates_declined_to_stepwise_kernel": stats[1], "residual_of_next_sweep": res,
    }
    if not args.no_cpu_baseline:
        from oracle import apply_oracle as A

        with pkg.BPXContext(0) as ctx:  # a fresh state for the oracle sample (the timed one has been evolved)
            problems.upload(ctx, p)
            ctx.sweep(3, 0.0, True)
            msgs0 = ctx.get_messages()
        link = lambda v, w: ("l", min(v, w), max(v, w))
        state = {}
        for v in range(ga.nv):
            nb = [ga.dst[e] for e in range(ga.row_ptr[v], ga.row_ptr[v + 1])]
            state[v] = (np.asarray(p.tensors[v]), (("s", v),) + tuple(link(v, w) for w in nb))
        env = {(ga.src[e], ga.dst[e]): msgs0[e] for e in range(ga.ne)}
        es = layers[0][len(layers[0]) // 2:][:2]
        ops = ops_for(len(es))
        t0 = time.perf_counter()
        for e, op in zip(es, ops):
            names = (("s", ga.src[e]), ("s", ga.dst[e]))
            A.apply_operator((op, names, names), state, env, trunc=chi, normalize=True)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": len(es) / dt, "unit": "gates/s", "cores": 1, "kind": "port",
                               "sample": f"{len(es)} bulk gates of layer 0, numpy oracle of apply_operators.jl:246-283 (oracle/apply_oracle.py)"}
    print(json.dumps(out))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "apply":
        if world > 1:
            raise SystemExit("--workload apply is a single-GPU workload")
        run_apply(args)
    elif args.single_process and args.gpus > 1 and world == 1:
        from tools import bench_single_process  # the single-process, multi-device variant (bpx_create_multi)

        bench_single_process.run(args)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
