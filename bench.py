#!/usr/bin/env python
"""bench.py -- headline benchmark of the Wigner-3j hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (reference-shaped oracle)

A *step* is one pass of the hot path over the BASELINE workload at lmax = 6143:
TT MCM + fused M++/M-- (EE/BB) MCM from two distinct masks, and the TTTT / EEEE / TETE coupled
covariance blocks (4 masks, product-mask window spectra).  Work unit = one 3j term, counted
over the FULL families exactly as the reference evaluates them:
7 reference families x T_fam(6143) = 5.41e11 terms per step.

Under torchrun (N > 1) the l1 rows are split into N work-balanced bands, every rank computes
its band, slabs are gathered to rank 0 with NCCL send/recv, rank 0 fills both triangles.
Total work is fixed => strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3j_terms_per_s (TT+EE/BB MCM and TTTT/EEEE/TETE coupledcov, lmax=6143)"
UNIT = "terms/s"
NOMINAL_FP64_TFLOPS = 37.2        # 148 SM x 64 FP64 lanes x 2 x 1.965 GHz

# job name -> (api, code, reference families, [(families fed, n_acc)] for the declared flop model
#              F = 20 + 2 n_acc per term)
JOBS = [
    ("M00", "mcm", 0, 1, [(1, 1)]),
    ("Mpp_Mmm", "mcm", 4, 2, [(2, 1)]),
    ("TTTT", "cov", 0, 1, [(1, 8)]),
    ("EEEE", "cov", 1, 1, [(1, 8)]),
    ("TETE", "cov", 3, 2, [(1, 4), (1, 1)]),
]


# From profiles/r01_ncu_final_summary.txt (ncu --set full, 1 GPU, lmax 6143): per-launch DRAM bytes and
# FP64 pipe activity of each pair kernel.  Quoted next to the live numbers, never used to compute them.
NCU_R01 = {
    "M00": {"dram_bytes": 94.4e6, "fp64_pipe_pct": 80.6}, "Mpp_Mmm": {"dram_bytes": 255.4e6, "fp64_pipe_pct": 82.1},
    "TTTT": {"dram_bytes": 97.0e6, "fp64_pipe_pct": 72.5}, "EEEE": {"dram_bytes": 98.0e6, "fp64_pipe_pct": 81.6},
    "TETE": {"dram_bytes": 97.7e6, "fp64_pipe_pct": 79.3},
}


def job_flops_per_tfam(job):
    return sum(f * (20 + 2 * n) for f, n in job[4])


def t_fam(lmax, lo=0, hi=None):
    hi = lmax + 1 if hi is None else hi
    l = np.arange(lo, hi, dtype=np.int64)
    return int(np.sum((2 * l + 1) * (lmax - l + 1)))


# ------------------------------------------------------------------------------------------
# inputs (host, numpy) -- synthetic zonal apodised masks, analytic spectra
# ------------------------------------------------------------------------------------------
def make_inputs(lmax):
    import powerspectra_jl_b200 as ps
    from powerspectra_jl_b200 import synthetic as syn
    V = syn.mask_spectra(lmax, seeds=(1001, 1002))
    ws, sp, rt = syn.covariance_inputs(lmax)
    i, j, p, q = ws.field_names
    N_ = ps.covariance.NULL
    W = lambda *k: ps.window_function_W(ws, *k).parent
    inp = {
        "M00": dict(V=V[(0, 1)]),
        "Mpp_Mmm": dict(V=V[(1, 1)]),
        "TTTT": dict(
            sp=[sp["TT", i, p].parent, sp["TT", j, q].parent, sp["TT", i, q].parent, sp["TT", j, p].parent],
            rt=[rt["TT", i, p].parent, rt["TT", j, q].parent, rt["TT", i, q].parent, rt["TT", j, p].parent],
            W=[W(N_, N_, i, p, "TT", j, q, "TT"), W(N_, N_, i, q, "TT", j, p, "TT"),
               W(N_, "TT", i, p, "TT", j, q, "TT"), W(N_, "TT", j, q, "TT", i, p, "TT"),
               W(N_, "TT", i, q, "TT", j, p, "TT"), W(N_, "TT", j, p, "TT", i, q, "TT"),
               W("TT", "TT", i, p, "TT", j, q, "TT"), W("TT", "TT", i, q, "TT", j, p, "TT")]),
        "EEEE": dict(
            sp=[sp["EE", i, p].parent, sp["EE", j, q].parent, sp["EE", i, q].parent, sp["EE", j, p].parent],
            rt=[rt["EE", i, p].parent, rt["EE", j, q].parent, rt["EE", i, q].parent, rt["EE", j, p].parent],
            W=[W(N_, N_, i, p, "PP", j, q, "PP"), W(N_, N_, i, q, "PP", j, p, "PP"),
               W(N_, "PP", i, p, "PP", j, q, "PP"), W(N_, "PP", j, q, "PP", i, p, "PP"),
               W(N_, "PP", i, q, "PP", j, p, "PP"), W(N_, "PP", j, p, "PP", i, q, "PP"),
               W("PP", "PP", i, p, "PP", j, q, "PP"), W("PP", "PP", i, q, "PP", j, p, "PP")]),
        "TETE": dict(
            sp=[sp["TT", i, p].parent, sp["EE", j, q].parent, sp["TE", i, q].parent, sp["TE", j, p].parent],
            rt=[rt["TT", i, p].parent, rt["EE", j, q].parent],
            W=[W(N_, N_, i, p, "TT", j, q, "PP"), W(N_, N_, i, q, "TP", j, p, "PT"),
               W(N_, "PP", i, p, "TT", j, q, "PP"), W(N_, "TT", j, q, "PP", i, p, "TT"),
               W("TT", "PP", i, p, "TT", j, q, "PP")]),
    }
    return inp


# ------------------------------------------------------------------------------------------
# CPU arm: the reference-shaped oracle (C/OpenMP restatement; Julia is not installed anywhere)
# ------------------------------------------------------------------------------------------
def cpu_sample(inp, lmax, rstep, threads=None):
    """One bounded sample of the step on the CPU: rows l1 = rstep//2, +rstep, ... of every job.
    Returns (terms evaluated, seconds)."""
    from oracle import psoracle as po
    row0 = rstep // 2
    t0 = time.perf_counter()
    terms = 0
    for name, api, code, fam, _ in JOBS:
        a = inp[name]
        if api == "mcm":
            if code == 4:      # the reference evaluates the (0,-2,2) family once per block
                for k in (2, 3):
                    _, t = po.mcm(k, 0, lmax, a["V"], row0=row0, rstep=rstep, threads=threads, return_terms=True)
                    terms += t
            else:
                _, t = po.mcm(code, 0, lmax, a["V"], row0=row0, rstep=rstep, threads=threads, return_terms=True)
                terms += t
        else:
            _, t = po.cov(code, 0, lmax, a["sp"], a["rt"], a["W"], row0=row0, rstep=rstep, threads=threads,
                          return_terms=True)
            terms += t
    return terms, time.perf_counter() - t0


def pick_rstep(inp, lmax, target_s):
    """Choose the row stride so one sample costs ~target_s: a sparse probe first (few rows, so the
    threads are badly balanced and the rate is pessimistic), then a ~3 s probe, then the answer."""
    full_terms = sum(j[3] for j in JOBS) * t_fam(lmax)
    rstep = 1024 if lmax >= 4096 else 64
    cpu_sample(inp, lmax, 4 * rstep)              # warm the threads / page in the library
    rate = None
    for goal in (3.0, target_s):
        terms, dt = cpu_sample(inp, lmax, rstep)
        rate = terms / dt
        rstep = max(4, int(round(full_terms / (rate * goal))))
    return rstep, rate


def run_reference(args, lmax):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import psoracle as po
    po.build()
    inp = make_inputs(lmax)
    cores = po.max_threads()
    rstep, _ = pick_rstep(inp, lmax, target_s=max(4.0, min(20.0, 150.0 / (args.steps + args.warmup))))
    for _ in range(args.warmup):
        cpu_sample(inp, lmax, rstep)
    tt, tn = 0.0, 0
    for _ in range(args.steps):
        n, dt = cpu_sample(inp, lmax, rstep)
        tt += dt
        tn += n
    value = tn / tt
    full_terms = sum(j[3] for j in JOBS) * t_fam(lmax)
    sample = f"every {rstep}th l1 row (from row {rstep // 2}) of each of the 5 calls, {tn // args.steps:.3e} terms per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * full_terms / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lmax={lmax}: MCM TT + EE/BB(M++,M--), coupledcov TTTT+EEEE+TETE",
                   "note": "C/OpenMP restatement of the reference CPU path (oracle/psoracle.c), not Julia; "
                           "ms_per_step is the full-step time extrapolated from the row sample by exact term count"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


def run_gpu(args, lmax):
    import torch
    import torch.distributed as dist

    import powerspectra_jl_b200 as ps
    from powerspectra_jl_b200 import device as dev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")     # host-side waits that must not occupy the GPUs
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    N = lmax + 1
    inp = make_inputs(lmax)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = {k: {kk: (pin(v) if isinstance(v, np.ndarray) else [pin(x) for x in v]) for kk, v in d.items()}
            for k, d in inp.items()}
    to_dev = lambda d: {kk: (v.cuda(non_blocking=True) if torch.is_tensor(v) else [x.cuda(non_blocking=True) for x in v])
                        for kk, v in d.items()}
    edges = dev.band_edges(0, lmax, world)
    lo, hi = edges[rank], edges[rank + 1]

    # output buffers: full matrix on every rank (N^2 x 8 B = 302 MB each at lmax 6143)
    outs = {}
    for name, api, code, fam, _ in JOBS:
        outs[name] = [torch.empty((N, N), dtype=torch.float64, device="cuda") for _ in range(2 if code == 4 and api == "mcm" else 1)]
    h2d_bytes = sum(v.numel() * 8 if torch.is_tensor(v) else sum(x.numel() * 8 for x in v)
                    for d in host.values() for v in d.values())
    d2h_bytes = sum(len(v) for v in outs.values()) * N * N * 8
    host_out = None
    launches = {"n": 0}
    kernel_events = []          # (job name, start, end) of the pair kernels of the timed steps

    comm = torch.cuda.Stream()          # gather + finish (+ D2H) of job k overlap the pair kernel of job k+1

    def compute(dinp, record, host_dst=None):
        main = torch.cuda.current_stream()
        for name, api, code, fam, _ in JOBS:
            a = dinp[name]
            X = outs[name]
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if api == "mcm":
                dev.mcm_slab(code, 0, lmax, a["V"], X[0], X[1] if len(X) > 1 else None, lo, hi)
            else:
                dev.cov_slab(code, 0, lmax, a["sp"], a["rt"], a["W"], X[0], lo, hi)
            launches["n"] += 2          # v2_prep_w + pair_kernel_v2
            if record:
                e1.record()
                kernel_events.append((name, e0, e1))
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                for k, Xo in enumerate(X):
                    dev.gather_bands(Xo, edges, 0, rank, world)
                    if rank == 0:
                        dev.finish(Xo, 0, lmax, api == "mcm")
                        launches["n"] += 1
                        if host_dst is not None:
                            host_dst[name][k].copy_(Xo, non_blocking=True)
        main.wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- kernel-resident measurement: inputs already in HBM ----
    dinp = {k: to_dev(d) for k, d in host.items()}
    for _ in range(args.warmup):
        compute(dinp, False)
    launches["n"] = 0
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(lambda: compute(dinp, True), args.steps)
    n_launch = launches["n"]
    clocks = sampler.stop() if sampler else None
    torch.cuda.synchronize()
    per_job = {}
    for name, e0, e1 in kernel_events:
        per_job.setdefault(name, []).append(e0.elapsed_time(e1))

    # ---- end to end: host buffers in, host buffers out, through the public entry points ----
    # The reference-facing call is the C ABI (psb200_mcm / psb200_cov with HOST buffers and ngpus = N):
    # one process drives the N GPUs, each GPU copies its own band of the result to the host.  Under
    # torchrun rank 0 makes that call while the other ranks wait at a barrier (their GPUs are idle).
    # The one-process-per-GPU driver (H2D -> band kernels -> NCCL gather -> finish -> D2H on rank 0)
    # is timed as well and reported as e2e.per_rank_driver.
    L = ps.lib()
    DP = ps._lib.DP
    ms_driver = None
    if world > 1:
        host_out0 = {name: [torch.empty((N, N), dtype=torch.float64).pin_memory() for _ in v]
                     for name, v in outs.items()} if rank == 0 else None

        def driver_step():
            d = {k: to_dev(dd) for k, dd in host.items()}
            compute(d, False, host_dst=host_out0)
        driver_step()
        ms_driver = timed(driver_step, max(1, min(args.steps, 3))) / max(1, min(args.steps, 3))
        del host_out0

    e2e_steps = max(1, min(args.steps, 3))
    wall_e2e = 0.0
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
    if rank == 0:
        host_out = {name: [torch.empty((N, N), dtype=torch.float64).pin_memory().numpy() for _ in v]
                    for name, v in outs.items()}

        def ptrs(arrs):
            return (DP * max(len(arrs), 1))(*[a.ctypes.data_as(DP) for a in arrs])

        def e2e_step():
            for name, api, code, fam, _ in JOBS:
                a = inp[name]
                O = host_out[name]
                if api == "mcm":
                    rc = L.psb200_mcm(code, 0, lmax, a["V"].ctypes.data_as(DP), a["V"].size, O[0].ctypes.data_as(DP), N,
                                      O[1].ctypes.data_as(DP) if len(O) > 1 else None, world)
                else:
                    rc = L.psb200_cov(code, 0, lmax, ptrs(a["sp"]), len(a["sp"]), ptrs(a["rt"]), len(a["rt"]),
                                      ptrs(a["W"]), len(a["W"]), a["W"][0].size, O[0].ctypes.data_as(DP), N, world)
                ps._lib.check(rc)
        torch.cuda.synchronize()
        e2e_step()                                   # warm-up (allocations inside the library)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()                               # blocking host calls: wall clock brackets them
        wall_e2e = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier(group=cpu_group)               # gloo: the waiting ranks leave their GPUs idle
    ms_e2e = wall_e2e

    # per-rank sum of pair-kernel time (band balance evidence)
    my_ms = torch.tensor([sum(float(np.mean(v)) for v in per_job.values())], dtype=torch.float64, device="cuda")
    all_ms = [torch.zeros_like(my_ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_ms, my_ms)
    else:
        all_ms = [my_ms]
    per_rank_ms = [round(float(x.item()), 3) for x in all_ms]

    if rank == 0:
        terms_step = sum(j[3] for j in JOBS) * t_fam(lmax)
        ms_step = ms_total / args.steps
        value = terms_step / (ms_step * 1e-3)
        e2e_value = terms_step / (ms_e2e / e2e_steps * 1e-3)
        # dominant kernel + roofline (FP64 pipe): declared flops of this rank's band / mean launch time
        mean_ms = {k: float(np.mean(v)) for k, v in per_job.items()}
        dom = max(mean_ms, key=mean_ms.get)
        dj = [j for j in JOBS if j[0] == dom][0]
        band_tfam = t_fam(lmax, lo, hi)
        flops = job_flops_per_tfam(dj) * band_tfam
        achieved = flops / (mean_ms[dom] * 1e-3) / 1e12
        peak = dev.dfma_peak(1 << 14) / 1e12
        all_flops = sum(job_flops_per_tfam(j) for j in JOBS) * band_tfam
        all_kernel_ms = sum(mean_ms.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lmax={lmax}: MCM TT + EE/BB(M++,M--), coupledcov TTTT+EEEE+TETE "
                                   f"(7 reference families x T_fam={t_fam(lmax):.4e} terms per step)",
                       "parallelism": f"l1 row bands x{world}, NCCL gather to rank 0" if world > 1 else "1 GPU",
                       "band_edges": edges, "pair_kernel_ms_per_rank": per_rank_ms,
                       "l2": "outputs (6 x N^2 x 8 B = 1.8 GB per step) exceed L2; inputs are O(lmax) vectors",
                       "kernel": os.environ.get("PSB200_KERNEL", "default")},
            "gpu_launches": n_launch,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / e2e_steps,
                    "path": f"psb200_mcm / psb200_cov C-ABI host calls with ngpus={world} (pageable inputs, pinned "
                            "host outputs; every GPU copies its own band of the result to the host)",
                    "per_rank_driver": None if ms_driver is None else {
                        "ms_per_step": ms_driver, "value": terms_step / (ms_driver * 1e-3),
                        "path": "pinned host -> H2D -> band kernels -> NCCL gather -> finish -> D2H on rank 0"}},
            "roofline": {"bound": "fp64", "kernel": f"pair kernel of job {dom}", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "psb200_dfma_peak DFMA microbenchmark measured in this run "
                                        f"(MEASURED_PEAKS.json has no FP64 entry; nominal {NOMINAL_FP64_TFLOPS})",
                         "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                         "flops_model": "F = 20 + 2 n_acc declared flops per 3j term over full families (SURVEY.md 8d)",
                         "fp64_pipe_active_pct_ncu": NCU_R01.get(dom, {}).get("fp64_pipe_pct"),
                         "frac_note": "frac uses the DECLARED flops of SURVEY.md 8d (full families, sqrt and divide per term); "
                                      "the kernel executes fewer and cheaper terms, so the kernel-quality figure is "
                                      "fp64_pipe_active_pct_ncu (share of cycles the FP64 pipe is busy, ncu capture)",
                         "traffic": NCU_R01.get(dom, {}).get("dram_bytes") if world == 1 else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full "
                                         "capture profiles/r01_ncu_final_summary.txt (1 GPU, lmax 6143); algorithmic HBM bytes = "
                                         "the 8 N^2/2 output bytes (151 MB), part of which is still in L2 at kernel end",
                         "all_kernels": {"declared_tflops": all_flops / (all_kernel_ms * 1e-3) / 1e12,
                                         "ms": mean_ms}},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import psoracle as po
            po.build()
            rstep, _ = pick_rstep(inp, lmax, target_s=15.0)
            n, dt = cpu_sample(inp, lmax, rstep)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": po.max_threads(), "kind": "port",
                                    "sample": f"every {rstep}th l1 row of each of the 5 calls ({n:.3e} terms, {dt:.1f} s); "
                                              "C/OpenMP restatement of the reference CPU path, not Julia"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lmax", type=int, default=6143)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.lmax)
    else:
        run_gpu(args, args.lmax)


if __name__ == "__main__":
    main()
