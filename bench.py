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

What the JSON line carries besides the contract keys:
  roofline.frac      executed FP64 instructions / (time x measured DFMA issue rate) of the dominant kernel:
                     executed pair-steps (psb200_job_stats: the library's own tiling) x FP64 instructions per
                     pair-step (counted in the SASS of the loaded libpsb200.so, tools/sass_fp64.py)
  roofline.useful_frac   same with the LIVE pair-steps only (dead lockstep slots excluded)
  roofline.declared_frac the SURVEY 8d declared-flop figure (full families, sqrt + divide per term); > 1 is
                     possible because the kernel executes fewer and cheaper terms than declared
  multi_gpu_check    the N-GPU host call, the 1-GPU host call and the NCCL-gather driver compared bit for bit
  parity             strict north-star statistics of sampled rows against the CPU oracle (Float64 and long double)
  known_answers      the TT matrix of the e2e host call against 50-digit entries (tests/golden/mcm_entries_mp.npz)
  e2e.pageable_outputs   the same step into pageable result arrays: the library's staged delivery and the CUDA runtime's own
  e2e.host_arrays    several GPUs on a multi-node host: one-node against NUMA-interleaved page-locked result arrays
  e2e.mirror_delivery    several GPUs: the step with PSB200_MIRROR=1 (half the DMA volume), measured in a subprocess
  extra              fused master call, TE at lmax 3071, QuickPol, the lmax 12287 sweep point, device-side decoupling,
                     W-spectrum production
"""
from __future__ import annotations

import argparse
import ctypes as C
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
PARITY_RSTEP = 128                # rows of the in-bench strict-parity statistics (every 128th from row 64)
CPU_RSTEP = 32                    # the CPU legs time every 32nd l1 row (from row 16) of each call: same rows everywhere

# job name -> (api, code, reference families, [(families fed, n_acc)] for the declared flop model
#              F = 20 + 2 n_acc per term)
JOBS = [
    ("M00", "mcm", 0, 1, [(1, 1)]),
    ("Mpp_Mmm", "mcm", 4, 2, [(2, 1)]),
    ("TTTT", "cov", 0, 1, [(1, 8)]),
    ("EEEE", "cov", 1, 1, [(1, 8)]),
    ("TETE", "cov", 3, 2, [(1, 4), (1, 1)]),
]
SPIN2_JOBS = {"Mpp_Mmm", "EEEE", "TETE"}     # rows l1 < 2 go through low_rows_kernel (one more launch on the band that holds them)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of each pair kernel, from the committed ncu --set full
# capture named below (1 GPU, lmax 6143).  Quoted as `roofline.traffic`, never used to compute anything.
NCU_TRAFFIC = {"file": "profiles/r01_ncu_final_summary.txt",
               "bytes": {"M00": 94.4e6, "Mpp_Mmm": 255.4e6, "TTTT": 97.0e6, "EEEE": 98.0e6, "TETE": 97.7e6}}
_ncu_json = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
if os.path.exists(_ncu_json):
    with open(_ncu_json) as _f:
        NCU_TRAFFIC = json.load(_f)


def workload(lmax):
    """config.workload -- the SAME string in both arms."""
    return (f"lmax={lmax}: MCM TT + EE/BB(M++,M--), coupledcov TTTT+EEEE+TETE "
            f"(7 reference families x T_fam={t_fam(lmax):.4e} terms per step)")


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
        "M02": dict(V=V[(0, 1)]),
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
def cpu_oracle():
    """The timed CPU build: -O3 -march=native, compiled on THIS machine, all host cores (torchrun exports
    OMP_NUM_THREADS=1 to its workers: the team size is forced, not inherited)."""
    from oracle import psoracle as po
    po.use_native_build()
    cores = po.set_threads(po.host_cores())
    return po, cores


def cpu_sample(po, inp, lmax, rstep=CPU_RSTEP, jobs=JOBS, ld=False):
    """One bounded sample of the step on the CPU: rows l1 = rstep//2, +rstep, ... of every job.
    Returns (terms evaluated, seconds, {job: [matrices]})."""
    row0 = rstep // 2
    t0 = time.perf_counter()
    terms = 0
    out = {}
    for name, api, code, fam, _ in jobs:
        a = inp[name]
        if api == "mcm":
            if code == 4:      # the reference evaluates the (0,-2,2) family once per block
                out[name] = []
                for k in (2, 3):
                    M, t = po.mcm(k, 0, lmax, a["V"], row0=row0, rstep=rstep, return_terms=True, ld=ld)
                    terms += t
                    out[name].append(M)
            else:
                M, t = po.mcm(code, 0, lmax, a["V"], row0=row0, rstep=rstep, return_terms=True, ld=ld)
                terms += t
                out[name] = [M]
        else:
            M, t = po.cov(code, 0, lmax, a["sp"], a["rt"], a["W"], row0=row0, rstep=rstep, return_terms=True, ld=ld)
            terms += t
            out[name] = [M]
    return terms, time.perf_counter() - t0, out


def sample_text(lmax, rstep, terms):
    nrows = len(range(rstep // 2, lmax + 1, rstep))
    return (f"every {rstep}th l1 row (rows {rstep // 2}, {rstep // 2 + rstep}, ...: {nrows} rows) of each of the 5 calls, "
            f"{terms:.3e} terms per sample; C/OpenMP restatement of the reference CPU path built -O3 -march=native "
            "on this host, not Julia")


def run_reference(args, lmax):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    po, cores = cpu_oracle()
    inp = make_inputs(lmax)
    for _ in range(args.warmup):
        cpu_sample(po, inp, lmax)
    tt, tn = 0.0, 0
    for _ in range(args.steps):
        n, dt, _ = cpu_sample(po, inp, lmax)
        tt += dt
        tn += n
    value = tn / tt
    full_terms = sum(j[3] for j in JOBS) * t_fam(lmax)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * full_terms / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload(lmax),
                   "note": "CPU arm: each step is the bounded row sample described in cpu_baseline.sample; ms_per_step is "
                           "the full-step time extrapolated from it by exact term count"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_text(lmax, CPU_RSTEP, tn / args.steps)},
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


def sass_counts(lib_path):
    """FP64 instructions per pair-step of every pair kernel, counted live in the loaded library's SASS
    (tools/sass_fp64.py); the committed count of the same build is the fallback when cuobjdump is absent."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_fp64
        res, _ = sass_fp64.analyse(lib_path)
        if res:
            return res, "cuobjdump -sass of the loaded libpsb200.so (tools/sass_fp64.py), this run"
    except Exception as e:          # noqa: BLE001 -- any failure falls back to the committed count
        sys.stderr.write(f"bench: live SASS count failed ({e}); using profiles/r02_sass_fp64.json\n")
    with open(os.path.join(ROOT, "profiles", "r02_sass_fp64.json")) as f:
        return json.load(f), "profiles/r02_sass_fp64.json (committed count of the same source)"


def job_stats(L, api, code, lmax, lenW, lo, hi):
    st = (C.c_longlong * 8)()
    rc = L.psb200_job_stats({"mcm": 0, "cov": 1, "master": 2}[api], code, lmax, lenW, lo, hi, st)
    assert rc == 0
    return {"exec_pair_steps": int(st[0]), "live_pair_steps": int(st[1]), "warps": int(st[2])}


def strict_stats(G, R, S=None, floor=1e-30):
    """North-star statistics of one sampled comparison: over every entry with |ref| > floor * max|row|,
    the largest relative error, the share above 1e-10 (ppm), and the same restricted to entries whose
    l3 sum cancels by less than 1e3 (S = condition sums, when given)."""
    rowmax = np.max(np.abs(R), axis=1, keepdims=True)
    sel = np.abs(R) > floor * rowmax
    if not sel.any():
        return {"n": 0}
    rel = np.abs(G[sel] - R[sel]) / np.abs(R[sel])
    out = {"n": int(sel.sum()), "strict_max_rel": float(rel.max()), "strict_fail_ppm": float(1e6 * np.mean(rel > 1e-10))}
    if S is not None:
        well = np.abs(S[sel]) <= 1e3 * np.abs(R[sel])
        out["well_conditioned_max_rel"] = float(rel[well].max()) if well.any() else 0.0
        out["err_over_condition_bound_max"] = float(np.max(np.abs(G[sel] - R[sel]) / (1e-10 * np.abs(R[sel]) + 1e-13 * np.abs(S[sel]))))
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
            for k, d in inp.items() if k in {j[0] for j in JOBS}}
    to_dev = lambda d: {kk: (v.cuda(non_blocking=True) if torch.is_tensor(v) else [x.cuda(non_blocking=True) for x in v])
                        for kk, v in d.items()}
    # folded split: rank r owns pieces r and 2 world - 1 - r of a 2 world-way kernel-cost split, launched as one tile list
    # (PSB200_BENCH_SPLIT=contiguous: one contiguous band per rank, the round-1 scheme)
    if os.environ.get("PSB200_BENCH_SPLIT") == "contiguous" and world > 1:
        e1 = dev.band_edges(0, lmax, world)
        owners = [[(e1[r], e1[r + 1])] for r in range(world)]
    else:
        owners = dev.folded_bands(0, lmax, world)
    mine = owners[rank]
    edges = [list(b) for bands in owners for b in bands]      # reported in config.band_edges

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

    gathered = {}                       # job name -> event: gather + finish (+ D2H) of its previous pass are done

    def compute(dinp, record, host_dst=None):
        """One step.  The pair kernel of job k+1 overlaps gather + finish (+ D2H) of job k on the side stream; a job's
        kernel waits only for the previous pass over ITS OWN output buffers, so consecutive steps overlap as well
        (everything is drained by the barrier + synchronize that closes a timed region)."""
        main = torch.cuda.current_stream()
        for name, api, code, fam, _ in JOBS:
            a = dinp[name]
            X = outs[name]
            if name in gathered:
                main.wait_event(gathered[name])
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if api == "mcm":
                dev.mcm_slab(code, 0, lmax, a["V"], X[0], X[1] if len(X) > 1 else None, bands=mine)
            else:
                dev.cov_slab(code, 0, lmax, a["sp"], a["rt"], a["W"], X[0], bands=mine)
            launches["n"] += 2 + (1 if (name in SPIN2_JOBS and mine[0][0] < 2) else 0)   # v3_prep_w + pair_kernel_v4 (+ low_rows_kernel)
            if record:
                e1.record()
                kernel_events.append((name, e0, e1))
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                for k, Xo in enumerate(X):
                    dev.gather_slabs(Xo, owners, 0, rank)
                    if rank == 0:
                        dev.finish(Xo, 0, lmax, api == "mcm")
                        launches["n"] += 1
                        if host_dst is not None:
                            host_dst[name][k].copy_(Xo, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(comm)
                gathered[name] = ev

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
    host_out0 = None
    if world > 1:
        host_out0 = {name: [torch.empty((N, N), dtype=torch.float64).pin_memory() for _ in v]
                     for name, v in outs.items()} if rank == 0 else None

        def driver_step():
            d = {k: to_dev(dd) for k, dd in host.items()}
            compute(d, False, host_dst=host_out0)
        driver_step()
        ms_driver = timed(driver_step, max(1, min(args.steps, 3))) / max(1, min(args.steps, 3))

    e2e_steps = max(1, min(args.steps, 3))
    wall_e2e = 0.0
    check = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)

    def ptrs(arrs):
        return (DP * max(len(arrs), 1))(*[a.ctypes.data_as(DP) for a in arrs])

    def host_calls(dst, ngpus, jobs=JOBS):
        for name, api, code, fam, _ in jobs:
            a = inp[name]
            O = dst[name]
            if api == "mcm":
                rc = L.psb200_mcm(code, 0, lmax, a["V"].ctypes.data_as(DP), a["V"].size, O[0].ctypes.data_as(DP), N,
                                  O[1].ctypes.data_as(DP) if len(O) > 1 else None, ngpus)
            else:
                rc = L.psb200_cov(code, 0, lmax, ptrs(a["sp"]), len(a["sp"]), ptrs(a["rt"]), len(a["rt"]),
                                  ptrs(a["W"]), len(a["W"]), a["W"][0].size, O[0].ctypes.data_as(DP), N, ngpus)
            ps._lib.check(rc)

    if rank == 0:
        host_out = {name: [torch.empty((N, N), dtype=torch.float64).pin_memory().numpy() for _ in v]
                    for name, v in outs.items()}
        torch.cuda.synchronize()
        host_calls(host_out, world)                  # warm-up (allocations inside the library)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_calls(host_out, world)              # blocking host calls: wall clock brackets them
        wall_e2e = (time.perf_counter() - t0) * 1e3
        e2e_alloc = {"used": "one-node page-locked arrays (torch pin_memory)", "numa_nodes": int(L.psb200_host_numa_nodes())}
        # Several GPUs on a multi-socket host: the same calls into result arrays whose 2 MB pieces alternate between the
        # NUMA nodes (psb200_host_alloc policy 1) -- every GPU then writes half of its region to its own socket.  Both
        # allocations are timed in every such run and both numbers are reported; e2e is the faster one.
        if world > 1 and e2e_alloc["numa_nodes"] > 1 and not os.environ.get("PSB200_BENCH_NO_INTERLEAVE"):
            import ctypes
            blocks = []
            try:                                        # an optional second measurement: it must never cost the line
                host_il = {}
                for name, v in outs.items():
                    host_il[name] = []
                    for _ in v:
                        p = L.psb200_host_alloc(N * N * 8, 1)
                        if not p:
                            raise RuntimeError("psb200_host_alloc: " + L.psb200_last_error().decode())
                        blocks.append(p)
                        host_il[name].append(np.ctypeslib.as_array(ctypes.cast(p, DP), shape=(N, N)))
                cnt = (ctypes.c_int * 8)()
                seen = L.psb200_host_placement(blocks[0], cnt, 8)
                host_calls(host_il, world)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    host_calls(host_il, world)
                wall_il = (time.perf_counter() - t0) * 1e3
                same = all(np.array_equal(a, b) for name in host_out for a, b in zip(host_out[name], host_il[name]))
                e2e_alloc.update({"one_node_ms_per_step": wall_e2e / e2e_steps, "interleaved_ms_per_step": wall_il / e2e_steps,
                                  "interleaved_placement_sample": list(cnt)[:e2e_alloc["numa_nodes"]] if seen > 0 else None,
                                  "interleaved_equals_one_node": bool(same)})
                if same and wall_il < wall_e2e:
                    wall_e2e = wall_il
                    e2e_alloc["used"] = "page-locked arrays interleaved over the NUMA nodes (psb200_host_alloc policy 1)"
                del host_il
            except Exception as exc:                    # reported, not fatal: e2e stays the one-node measurement
                e2e_alloc["interleaved_error"] = repr(exc)
            for p in blocks:
                L.psb200_host_free(p)

        # ---- mirror delivery (PSB200_MIRROR=1, off by default): only the block columns of the results cross PCIe and the
        # library's host threads write the symmetric side from the same bytes -- bit-identical by construction (the same
        # IEEE products), half the DMA volume.  Timed in every run next to the standard delivery; e2e is the faster of
        # the two when, and only when, the matrices are identical. ----
        mirror = None
        if world == 1:
            mirror = {"skipped": "one GPU: the host threads, not PCIe, bound the mirror delivery there "
                                 "(profiles/r02_mirror_probe_n1.jsonl: identical matrices, 10.5 vs 6.7 ms for TT)"}
        elif not os.environ.get("PSB200_BENCH_NO_MIRROR"):
            # in a process of its own (tools/e2e_step_probe.py: the same five host calls, standard and mirror delivery,
            # page-locked arrays): an optional path must never be able to cost the line
            import subprocess
            try:
                env = {k: v for k, v in os.environ.items() if k not in ("PSB200_MIRROR", "OMP_NUM_THREADS")}
                out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "e2e_step_probe.py"),
                                      str(world), str(e2e_steps), str(lmax)], capture_output=True, text=True, timeout=300, env=env)
                mirror = json.loads(out.stdout.strip().splitlines()[-1])
                mirror["measured_in"] = "subprocess tools/e2e_step_probe.py (same box, same GPUs, the ranks idle)"
                mirror["in_process_standard_ms_per_step"] = wall_e2e / e2e_steps
                mirror["used_for_e2e"] = bool(mirror["equals_standard_delivery"]
                                              and mirror["mirror_ms_per_step"] < min(mirror["standard_ms_per_step"], wall_e2e / e2e_steps))
                # what crosses PCIe with the mirror delivery: the block columns (lower triangles + the diagonal blocks of
                # the sub-bands, counted here as N (N + 1) / 2 entries per matrix)
                mirror["d2h_bytes_per_step"] = int(sum(len(v) for v in outs.values()) * (N * (N + 1) // 2) * 8)
                if mirror["used_for_e2e"]:
                    wall_e2e = mirror["mirror_ms_per_step"] * e2e_steps
                    d2h_bytes = mirror["d2h_bytes_per_step"]
            except Exception as exc:
                mirror = {"error": repr(exc)}

        # ---- the same calls into PAGEABLE result arrays, which is what the reference allocates (spectralzeros): the
        # library's staged delivery (page-locked ring + scatter threads) beside the CUDA runtime's own bounce copies ----
        pageable = None
        if not os.environ.get("PSB200_BENCH_NO_PAGEABLE"):
            saved = os.environ.get("PSB200_STAGED")
            try:                                        # an additional measurement: it must never cost the line
                host_pg = {name: [np.zeros((N, N)) + 0.0 for _ in v] for name, v in outs.items()}  # touched pages, like Julia's zeros
                pageable = {}
                for key, val in (("staged_ms_per_step", "1"), ("runtime_bounce_ms_per_step", "0")):
                    os.environ["PSB200_STAGED"] = val
                    host_calls(host_pg, world)
                    t0 = time.perf_counter()
                    for _ in range(e2e_steps):
                        host_calls(host_pg, world)
                    pageable[key] = (time.perf_counter() - t0) * 1e3 / e2e_steps
                    pageable[key.replace("_ms_per_step", "_equals_page_locked")] = bool(
                        all(np.array_equal(a, b) for name in host_out for a, b in zip(host_out[name], host_pg[name])))
                del host_pg
            except Exception as exc:
                pageable = {"error": repr(exc)}
            if saved is None:
                os.environ.pop("PSB200_STAGED", None)
            else:
                os.environ["PSB200_STAGED"] = saved
            if pageable.get("staged_equals_page_locked") is False or pageable.get("runtime_bounce_equals_page_locked") is False:
                raise SystemExit(f"bench: results in pageable arrays differ from the page-locked ones: {pageable}")

        # ---- the TT matrix of this very step against multiprecision known answers (tests/golden/mcm_entries_mp.npz: 96
        # entries at lmax 6143 computed with mpmath at 50 digits, independent of oracle/ and of the library) ----
        known = None
        try:
            gpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "mcm_entries_mp.npz")
            if lmax == 6143 and os.path.exists(gpath):
                g = np.load(gpath)
                if np.array_equal(g["V_6143"], inp["M00"]["V"]):
                    A = host_out["M00"][0]                       # A[l2, l1] = M[l1, l2] (column-major result in a C-ordered array)
                    worst, n = 0.0, 0
                    for (l1, l2), xi, sa in zip(g["pairs_6143"], g["xi_6143"][:, 0], g["sabs_6143"][:, 0]):
                        for i, j in ((l1, l2), (l2, l1)):
                            if (2 * j + 1) * sa <= 1e3 * abs((2 * j + 1) * xi):      # sums that do not cancel: strict criterion
                                worst = max(worst, abs(A[j, i] - (2 * j + 1) * xi) / abs((2 * j + 1) * xi))
                                n += 1
                    known = {"matrix": "M00 of the e2e host call", "entries": n, "strict_max_rel": worst, "north_star": 1e-10,
                             "reference": "mpmath, 50 digits (tests/golden/make_golden_highl.py)"}
                else:
                    known = {"skipped": "this host's numpy rounds the synthetic window spectrum differently from the fixture"}
        except Exception as exc:
            known = {"error": repr(exc)}
        if known and known.get("strict_max_rel", 0.0) > 1e-10:
            raise SystemExit(f"bench: the TT matrix misses the multiprecision known answers: {known}")

        # ---- the outputs themselves: N-GPU host call == 1-GPU host call == NCCL-gather driver, bit for bit ----
        # (every (l1,l2) pair is computed independently of the banding, src/modecoupling.jl:84-92, so any difference
        # is a bug in the band delivery / gather / finish plumbing)
        check = {"bitwise_equal": True, "max_abs_diff": 0.0, "matrices": 0, "compared": []}

        def cmp(tag, a, b):
            same = bool(np.array_equal(a, b))
            check["matrices"] += 1
            if not same:
                check["bitwise_equal"] = False
                check["max_abs_diff"] = max(check["max_abs_diff"], float(np.nanmax(np.abs(a - b))))
                check.setdefault("mismatch", []).append(tag)
        one = np.empty((N, N))
        one2 = np.empty((N, N))
        for job in JOBS:
            name = job[0]
            for k, Xo in enumerate(outs[name]):       # resident path of the timed steps (after gather + finish)
                cmp(f"{name}[{k}] resident driver vs host call", Xo.cpu().numpy(), host_out[name][k])
            if world > 1:
                for k in range(len(outs[name])):      # per-rank driver with its own D2H
                    cmp(f"{name}[{k}] per-rank driver D2H vs host call", host_out0[name][k].numpy(), host_out[name][k])
                host_calls({name: [one, one2]}, 1, jobs=[job])
                for k, ref1 in enumerate([one, one2][:len(outs[name])]):
                    cmp(f"{name}[{k}] ngpus=1 vs ngpus={world}", ref1, host_out[name][k])
        check["compared"] = (["device-API driver (bands + NCCL gather + finish) vs C-ABI host call"]
                             + ([f"C-ABI host call ngpus=1 vs ngpus={world}", "per-rank driver D2H vs C-ABI host call"] if world > 1 else []))
        del one, one2
    host_out0 = None
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
        # ---- roofline (FP64 pipe) of the dominant kernel: executed work, measured live ----
        mean_ms = {k: float(np.mean(v)) for k, v in per_job.items()}
        dom = max(mean_ms, key=mean_ms.get)
        peak = dev.dfma_peak(1 << 14) / 1e12                     # TFLOP/s, 2 flops per DFMA lane-instruction
        issue_peak = peak * 1e12 / 2.0                           # FP64 lane-instructions per second
        sass, sass_src = sass_counts(ps._lib.LIB_PATH)
        band_tfam = sum(t_fam(lmax, lo, hi) for lo, hi in mine)
        kern = {}
        for job in JOBS:
            name, api, code = job[0], job[1], job[2]
            lenW = inp[name]["V"].size if api == "mcm" else inp[name]["W"][0].size
            parts = [job_stats(L, api, code, lmax, lenW, lo, hi) for lo, hi in mine]
            st = {k: sum(p[k] for p in parts) for k in parts[0]}
            sc = sass[name]
            t = mean_ms[name] * 1e-3
            kern[name] = {
                "ms": mean_ms[name], "fp64_instr_per_pair_step": sc["fp64_per_pair_step"],
                "exec_pair_steps": st["exec_pair_steps"], "live_pair_steps": st["live_pair_steps"],
                "tiling_efficiency": st["live_pair_steps"] / max(1, st["exec_pair_steps"]),
                "frac": st["exec_pair_steps"] * sc["fp64_per_pair_step"] / t / issue_peak,
                "useful_frac": st["live_pair_steps"] * sc["fp64_per_pair_step"] / t / issue_peak,
                "executed_tflops": st["exec_pair_steps"] * sc["flops_per_pair_step"] / t / 1e12,
                "declared_frac": job_flops_per_tfam(job) * band_tfam / t / 1e12 / peak,
            }
        kd = kern[dom]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(lmax),
                       "parallelism": (f"l1 row bands x{world} (folded: a low and a high band per rank, one tile list), "
                                       "NCCL gather to rank 0") if world > 1 else "1 GPU",
                       "band_edges": edges, "pair_kernel_ms_per_rank": per_rank_ms,
                       "l2": "outputs (6 x N^2 x 8 B = 1.8 GB per step) exceed L2; inputs are O(lmax) vectors",
                       "kernel": os.environ.get("PSB200_KERNEL", "default")},
            "gpu_launches": n_launch,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / e2e_steps,
                    "path": f"psb200_mcm / psb200_cov C-ABI host calls with ngpus={world} (pageable inputs, pinned "
                            "host outputs; every GPU copies its own band of the result to the host"
                            + ("; mirror delivery, PSB200_MIRROR=1: see mirror_delivery)" if mirror and mirror.get("used_for_e2e") else ")"),
                    "host_arrays": e2e_alloc, "mirror_delivery": mirror, "pageable_outputs": pageable,
                    "per_rank_driver": None if ms_driver is None else {
                        "ms_per_step": ms_driver, "value": terms_step / (ms_driver * 1e-3),
                        "path": "pinned host -> H2D -> band kernels -> NCCL gather -> finish -> D2H on rank 0"}},
            "roofline": {"bound": "fp64", "kernel": f"pair kernel of job {dom} (rank 0 band)",
                         "achieved": kd["exec_pair_steps"] * kd["fp64_instr_per_pair_step"] / (kd["ms"] * 1e-3) / 1e12,
                         "peak": issue_peak / 1e12, "unit": "T FP64 lane-instr/s", "frac": kd["frac"],
                         "useful_frac": kd["useful_frac"], "declared_frac": kd["declared_frac"],
                         "how": "achieved = executed pair-steps (psb200_job_stats: lockstep steps x 32 R slots per warp, "
                                "dead slots included) x FP64 instructions per pair-step (DFMA+DMUL+DADD of the main loop in "
                                "the SASS) / mean CUDA-event time of the launch; peak = psb200_dfma_peak DFMA microbenchmark "
                                "of this run / 2 (MEASURED_PEAKS.json has no FP64 entry); useful_frac counts live pair-steps "
                                "only; declared_frac = SURVEY 8d flops (F = 20 + 2 n_acc per full-family term) / time / peak",
                         "sass_count_source": sass_src,
                         "peak_tflops_dfma": peak, "nominal_tflops": NOMINAL_FP64_TFLOPS,
                         "executed_tflops": kd["executed_tflops"],
                         "traffic": NCU_TRAFFIC["bytes"].get(dom) if world == 1 else None,
                         "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full "
                                         f"capture {NCU_TRAFFIC['file']} (1 GPU, lmax 6143); algorithmic HBM bytes = the 8 N^2/2 "
                                         "output bytes (151 MB), part of which is still in L2 at kernel end",
                         "all_kernels": kern},
            "multi_gpu_check": check,
            "known_answers": known,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            po, cores = cpu_oracle()
            n, dt, ref64 = cpu_sample(po, inp, lmax)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample_text(lmax, CPU_RSTEP, n) + f" ({dt:.1f} s)"}
            # strict north-star statistics on a sparser row sample: GPU and Float64 oracle against long double
            _, _, r64 = cpu_sample(po, inp, lmax, rstep=PARITY_RSTEP)
            _, _, rld = cpu_sample(po, inp, lmax, rstep=PARITY_RSTEP, ld=True)
            rows = np.arange(PARITY_RSTEP // 2, lmax + 1, PARITY_RSTEP)
            par = {}
            for name in r64:
                for k, (R64, RLD) in enumerate(zip(r64[name], rld[name])):
                    tag = name if len(r64[name]) == 1 else f"{name}[{k}]"
                    G = host_out[name][k]          # column-major (l1, l2) = numpy [l2, l1]
                    gm = np.zeros((len(rows), N)); om = np.zeros_like(gm); rm = np.zeros_like(gm)
                    for i, r in enumerate(rows):   # sampled row l1 = r, columns l2 >= l1
                        gm[i, r:] = G[r:, r]
                        om[i, r:] = R64[r, r:]
                        rm[i, r:] = RLD[r, r:]
                    par[tag] = {"gpu_vs_long_double": strict_stats(gm, rm), "f64_oracle_vs_long_double": strict_stats(om, rm),
                                "gpu_vs_f64_oracle": strict_stats(gm, om)}
            line["parity"] = {"rows": f"every {PARITY_RSTEP}th l1 row from {PARITY_RSTEP // 2}, columns l2 >= l1",
                              "criterion": "north star: |err|/|ref| <= 1e-10 on every entry above 1e-30 of its row maximum; "
                                           "entries whose l3 sum cancels by > 1e3 cannot meet it in ANY Float64 evaluation "
                                           "(see f64_oracle_vs_long_double); the test suite gates on the condition-aware bound",
                              "per_matrix": par}
        if not args.no_extra:
            try:
                line["extra"] = extras(args, ps, dev, L, world, torch, cpu=(world == 1 and not args.no_cpu))
            except Exception as e:      # noqa: BLE001 -- the extras never cost the headline line
                import traceback
                traceback.print_exc()
                line["extra"] = {"error": repr(e)}
        print(json.dumps(line))
        if check is not None and not check["bitwise_equal"]:
            sys.stderr.write("bench: multi_gpu_check FAILED: " + json.dumps(check) + "\n")
            if world > 1:
                dist.destroy_process_group()
            sys.exit(3)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# extras: measured outside the headline timed region, on rank 0's GPU(s) through the C ABI
# ------------------------------------------------------------------------------------------
def extras(args, ps, dev, L, world, torch, cpu):
    from powerspectra_jl_b200 import synthetic as syn
    DP = ps._lib.DP
    out = {}

    def dtime(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def wtime(fn, reps=2):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) * 1e3 / reps

    po = None
    if cpu:
        po, cores = cpu_oracle()

    # (1) fused master call vs the separate calls it replaces, lmax 6143, resident, 1 GPU
    lmax = args.lmax
    N = lmax + 1
    sky_V = syn.mask_spectra(lmax, seeds=(1001, 1002, 1003, 1004))
    # V_TT (T1 x T2), V_TP (T1 x P2), V_PT (P1 x T2), V_PP (P1 x P2) with T = masks 0, 2 and P = masks 1, 3
    Vd = {(0, 0): torch.tensor(sky_V[(0, 2)], device="cuda"), (0, 1): torch.tensor(sky_V[(0, 3)], device="cuda"),
          (1, 0): torch.tensor(sky_V[(1, 2)], device="cuda"), (1, 1): torch.tensor(sky_V[(1, 3)], device="cuda")}
    X = [torch.empty((N, N), dtype=torch.float64, device="cuda") for _ in range(5)]

    def fused():
        dev.master_slab(0, lmax, Vd[(0, 0)], Vd[(0, 1)], Vd[(1, 0)], Vd[(1, 1)], X)
        for k in range(5):
            dev.finish(X[k], 0, lmax, True)

    def separate():
        dev.mcm_slab(0, 0, lmax, Vd[(0, 0)], X[0])
        dev.mcm_slab(1, 0, lmax, Vd[(0, 1)], X[1])
        dev.mcm_slab(1, 0, lmax, Vd[(1, 0)], X[2])
        dev.mcm_slab(4, 0, lmax, Vd[(1, 1)], X[3], X[4])
        for k in range(5):
            dev.finish(X[k], 0, lmax, True)
    t_f, t_s = dtime(fused), dtime(separate)
    out["master_fused"] = {"lmax": lmax, "fused_ms": t_f, "separate_calls_ms": t_s, "speedup": t_s / t_f,
                           "terms_per_s": 7 * t_fam(lmax) / (t_f * 1e-3),
                           "what": "psb200_mcm_master_dev (M00, M02 x2, M++, M--) vs psb200_mcm_dev kinds 0, 1, 1, 4; device-resident, 1 GPU"}
    del X, Vd
    torch.cuda.empty_cache()

    # (2) TE mode-coupling matrix at lmax 3071 (BASELINE configs[1]) and (4) the lmax 12287 sweep point (configs[4])
    for tag, lm, kind, fam in (("TE_lmax3071", 3071, 1, 2), ("TT_lmax12287", 12287, 0, 1), ("EE_BB_lmax12287", 12287, 4, 2)):
        n = lm + 1
        V = syn.mask_spectra(lm, seeds=(1001, 1002))[(0, 1)]
        Vt = torch.tensor(V, device="cuda")
        nout = 2 if kind == 4 else 1
        Xs = [torch.empty((n, n), dtype=torch.float64, device="cuda") for _ in range(nout)]

        def resident():
            dev.mcm_slab(kind, 0, lm, Vt, Xs[0], Xs[1] if nout > 1 else None)
            for x in Xs:
                dev.finish(x, 0, lm, True)
        t_r = dtime(resident, reps=2)
        del Xs
        torch.cuda.empty_cache()
        Hs = [torch.empty((n, n), dtype=torch.float64).pin_memory().numpy() for _ in range(nout)]

        def e2e():
            ps._lib.check(L.psb200_mcm(kind, 0, lm, V.ctypes.data_as(DP), V.size, Hs[0].ctypes.data_as(DP), n,
                                       Hs[1].ctypes.data_as(DP) if nout > 1 else None, world))
        t_e = wtime(e2e, reps=2)
        terms = fam * t_fam(lm)
        rec = {"lmax": lm, "kind": kind, "ms_resident_1gpu": t_r, "terms_per_s": terms / (t_r * 1e-3),
               "e2e_ms": t_e, "e2e_terms_per_s": terms / (t_e * 1e-3), "e2e_ngpus": world,
               "d2h_bytes": nout * n * n * 8}
        if po is not None:
            rstep = 32 if lm <= 4096 else 128
            t0 = time.perf_counter()
            tn = 0
            for k in ((2, 3) if kind == 4 else (kind,)):
                _, t = po.mcm(k, 0, lm, V, row0=rstep // 2, rstep=rstep, return_terms=True)
                tn += t
            dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": tn / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"every {rstep}th l1 row from {rstep // 2} ({tn:.3e} terms, {dt:.1f} s)"}
            rec["speedup_vs_cpu_e2e"] = rec["e2e_terms_per_s"] / (tn / dt)
        out[tag] = rec
        del Hs

    # (3) QuickPol Xi matrix: lmax 6143, band +-128, (nu1, nu2, s1, s2) = (2, -2, 2, 2)
    lm, band = args.lmax, 128
    nb = 2 * band + 1
    rng = np.random.default_rng(7)
    l = np.arange(2 * lm + 1)
    W = rng.normal(size=l.size) / (1.0 + l / 40.0) ** 2 + 1.0 / (1.0 + l) ** 1.5
    nu1, nu2, s1, s2 = 2, -2, 2, 2
    qterms = 0
    for lpp in range(2, lm + 1):
        ls = np.arange(max(2, lpp - band), min(lm, lpp + band) + 1)
        d = np.abs(ls - lpp)
        for m1 in (s1 + nu1, s2 + nu2):
            qterms += int(np.sum(np.maximum(ls + lpp - np.maximum(d, abs(m1)) + 1, 0)))
    dW = torch.tensor(W, device="cuda")
    dX = torch.zeros((lm + 1, nb), device="cuda", dtype=torch.float64)
    t_r = dtime(lambda: dev.quickpol_slab(nu1, nu2, s1, s2, lm, dW, dX, band, band), reps=3)
    Xb = np.zeros((lm + 1, nb))

    def qe2e():
        ps._lib.check(L.psb200_quickpol_xi(nu1, nu2, s1, s2, lm, W.ctypes.data_as(DP), W.size, band, band,
                                           Xb.ctypes.data_as(DP), nb, world))
    t_e = wtime(qe2e, reps=2)
    rec = {"lmax": lm, "band": band, "case": [nu1, nu2, s1, s2], "terms": qterms, "ms_resident_1gpu": t_r,
           "terms_per_s": qterms / (t_r * 1e-3), "e2e_ms": t_e, "e2e_terms_per_s": qterms / (t_e * 1e-3), "e2e_ngpus": world}
    if po is not None:
        # CPU sample: the band of the first 1/16 of the rows is cheap and unrepresentative; time a smaller matrix of
        # the same band instead and quote its own term rate
        lm_c = 1535
        Wc = W[:2 * lm_c + 1]
        t0 = time.perf_counter()
        _, tn = po.quickpol_xi(nu1, nu2, s1, s2, lm_c, Wc, band, band, dense=False, return_terms=True)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": tn / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"the whole Xi matrix at lmax {lm_c}, same band ({tn:.3e} terms, {dt:.1f} s)"}
        rec["speedup_vs_cpu_e2e"] = rec["e2e_terms_per_s"] / (tn / dt)
    out["quickpol"] = rec

    # (5) decoupling on the device: everything maskedalm2spectra solves, matrices never leave the GPUs
    lmax = args.lmax
    N = lmax + 1 - 2
    Vs = [np.ascontiguousarray(sky_V[k]) for k in ((0, 2), (0, 3), (1, 2), (1, 3))]
    rng = np.random.default_rng(11)
    pcl = np.asfortranarray(rng.normal(size=(N, 9)))
    cl = np.zeros_like(pcl, order="F")

    def solve():
        ps._lib.check(L.psb200_master_solve(2, lmax, Vs[0].ctypes.data_as(DP), Vs[1].ctypes.data_as(DP), Vs[2].ctypes.data_as(DP),
                                            Vs[3].ctypes.data_as(DP), lmax + 1, pcl.ctypes.data_as(DP), N, cl.ctypes.data_as(DP), N, world))
    t_s = wtime(solve, reps=2)
    Hm = [torch.empty((N, N), dtype=torch.float64).pin_memory().numpy() for _ in range(5)]

    def deliver():
        ps._lib.check(L.psb200_mcm_master(2, lmax, Vs[0].ctypes.data_as(DP), Vs[1].ctypes.data_as(DP), Vs[2].ctypes.data_as(DP),
                                          Vs[3].ctypes.data_as(DP), lmax + 1, *[h.ctypes.data_as(DP) for h in Hm], N, world))
    t_d = wtime(deliver, reps=2)
    out["decouple_on_device"] = {
        "lmax": lmax, "lmin": 2, "ngpus": world, "master_solve_ms": t_s, "mcm_master_to_host_ms": t_d,
        "d2h_bytes_avoided": 5 * N * N * 8, "d2h_bytes_of_the_solve": 9 * N * 8,
        "what": "psb200_master_solve: fused five-matrix pass on all GPUs, bands stored into GPU 0 over NVLink, LU of M00, "
                "M02 x2 and the two dense 2N x 2N EE/BB, EB/BE block systems (cuSOLVER getrf/getrs), nine decoupled spectra "
                "back; beside it psb200_mcm_master delivering the five matrices to pinned host memory (before any host LU)"}
    del Hm
    n = 2048 if cpu else N + 2
    Y = np.asfortranarray(rng.normal(size=(n, n)))
    B1 = np.asfortranarray(rng.normal(size=(n, n)) + n ** 0.5 * np.eye(n))
    B2 = np.asfortranarray(rng.normal(size=(n, n)) + n ** 0.5 * np.eye(n))
    O = np.zeros((n, n), order="F")

    def dec():
        ps._lib.check(L.psb200_decouple_covmat(n, Y.ctypes.data_as(DP), n, B1.ctypes.data_as(DP), n, B2.ctypes.data_as(DP), n,
                                               O.ctypes.data_as(DP), n))
    t_c = wtime(dec, reps=2)
    rec = {"n": n, "device_ms": t_c, "what": "psb200_decouple_covmat (host buffers in and out, pageable): B1^-1 Y (B2^-1)^T"}
    if cpu:
        t0 = time.perf_counter()
        Hh = np.linalg.solve(B2, np.linalg.solve(B1, Y).T).T
        rec["host_lapack_ms"] = (time.perf_counter() - t0) * 1e3
        rec["max_rel_diff_vs_host"] = float(np.max(np.abs(O - Hh)) / np.max(np.abs(Hh)))
    out["decouple_covmat"] = rec

    # (6) W-spectrum production: what effective_weight_alm! / window_function_W! ask of Healpix.jl --
    # map2alm(mask_i .* mask_j .* sigma^2 .* Omega_pix; lmax, niter = 3) and alm2cl -- at the map size of the headline lmax
    nside = 1
    while 3 * nside < args.lmax + 1:
        nside *= 2
    lm = min(args.lmax, 3 * nside - 1)
    npix = 12 * nside * nside
    rng = np.random.default_rng(21)
    gen = torch.Generator(device="cuda").manual_seed(21)
    dalm = torch.randn(dev.alm_size(lm), dtype=torch.complex128, device="cuda", generator=gen)
    ls = torch.tensor(np.concatenate([np.arange(m, lm + 1) for m in range(lm + 1)]), device="cuda")
    dalm *= torch.exp(-0.5 * (ls / (0.25 * lm)) ** 2)
    dalm[:lm + 1] = dalm[:lm + 1].real.to(torch.complex128)
    del ls
    dmap = torch.empty(npix, dtype=torch.float64, device="cuda")
    dback = torch.empty_like(dalm)
    dcl = torch.empty(lm + 1, dtype=torch.float64, device="cuda")
    t_syn = dtime(lambda: dev.alm2map_dev(nside, lm, dalm, dmap), reps=2)
    t_ana = dtime(lambda: dev.map2alm_dev(nside, lm, dmap, dback, 0), reps=2)
    t_m2a = dtime(lambda: dev.map2alm_dev(nside, lm, dmap, dback, 3), reps=2)
    err3 = float((dback - dalm).abs().max() / dalm.abs().max())
    t_cl = dtime(lambda: dev.alm2cl_dev(lm, dback, dalm, dcl), reps=3)
    st = dev.sht_stats(nside, lm)
    peak = L.psb200_dfma_peak(20000) / 2.0
    hm = [np.ascontiguousarray(dmap.cpu().numpy()), np.full(npix, 0.5), np.full(npix, 2.0)]
    halm = np.zeros(dev.alm_size(lm), dtype=np.complex128)
    ptrs = (DP * 3)(*[h.ctypes.data_as(DP) for h in hm])

    def m2a_host():
        ps._lib.check(L.psb200_map2alm(nside, lm, 3, 3, ptrs, 4.0 * np.pi / npix, halm.ctypes.data_as(DP)))
    t_e = wtime(m2a_host, reps=1)
    out["w_production"] = {
        "nside": nside, "lmax": lm, "niter": 3, "ring_pairs_per_lane": st["R"],
        "alm2map_ms": t_syn, "analysis_ms": t_ana, "map2alm_niter3_ms": t_m2a, "alm2cl_ms": t_cl,
        "map2alm_e2e_ms": t_e, "e2e_h2d_bytes": 3 * npix * 8, "e2e_d2h_bytes": int(halm.nbytes),
        "roundtrip_rel_err_niter3": err3,
        "legendre_steps_per_pass": st["exec_steps"], "live_over_executed": st["live_steps"] / st["exec_steps"],
        "fp64_frac_analysis": 4.0 * st["exec_steps"] / (t_ana * 1e-3) / peak,
        "fp64_frac_synthesis": 4.0 * st["exec_steps"] / (t_syn * 1e-3) / peak,
        "fp64_frac_map2alm": 7 * 4.0 * st["exec_steps"] / (t_m2a * 1e-3) / peak,
        "what": "psb200_map2alm_dev / psb200_alm2map_dev / psb200_alm2cl_dev device-resident on one GPU; e2e = psb200_map2alm "
                "with three pageable host maps (mask_i, mask_j, sigma^2) in and the alm out; fractions = executed "
                "(l, m, ring pair) steps x 4 FP64 instructions / time of the whole pass (ring FFT stage included) / DFMA peak"}
    # the same through the batched call: four products of the three maps, every map uploaded once, products dealt to `world` GPUs
    prods = [[0, 1, -1], [0, 2, -1], [1, 2, -1], [0, 1, 2]]
    outs = [np.zeros(dev.alm_size(lm), dtype=np.complex128) for _ in prods]
    idx = (C.c_int * (3 * len(prods)))(*[v for p_ in prods for v in p_])
    sc = np.full(len(prods), 4.0 * np.pi / npix)
    op = (DP * len(prods))(*[o.ctypes.data_as(DP) for o in outs])

    def many():
        ps._lib.check(L.psb200_map2alm_many(nside, lm, 3, 3, ptrs, len(prods), idx, sc.ctypes.data_as(DP), op, world))
    t_b = wtime(many, reps=1)
    out["w_production"]["batched"] = {"products": len(prods), "unique_maps": 3, "ngpus": world, "ms_per_product": t_b / len(prods),
                                      "equals_single_call": bool(np.array_equal(outs[3], halm)),
                                      "what": "psb200_map2alm_many: maps uploaded once per device, alm downloads overlapped with the next product"}
    del dalm, dmap, dback, outs
    L.psb200_sht_release()
    if cpu:
        # CPU arm in the reference's shape (oracle/shtcpu.c: per-ring FFTs + scaled lambda_lm recurrences, plain C + OpenMP,
        # -O3 -march=native; NOT libsharp, which Healpix.jl calls and which is not in this image), at nside 512 where it
        # takes seconds; the GPU is timed at the same size beside it
        from oracle import shtoracle as so
        ns_c, lm_c = 512, 1535
        fc = np.random.default_rng(22).normal(size=12 * ns_c * ns_c)
        so.fast_lib(native=True).shtcpu_threads(cores)
        t0 = time.perf_counter()
        a_cpu = so.fast_map2alm(fc, ns_c, lm_c, 3, native=True)
        dt = time.perf_counter() - t0
        dfc = torch.tensor(fc, device="cuda")
        dac = torch.empty(dev.alm_size(lm_c), dtype=torch.complex128, device="cuda")
        t_g = dtime(lambda: dev.map2alm_dev(ns_c, lm_c, dfc, dac, 3), reps=3)
        out["w_production"]["cpu_baseline"] = {
            "value": 1.0 / dt, "unit": "map2alm(niter=3)/s at nside 512, lmax 1535", "cores": cores, "kind": "port",
            "sample": f"one whole transform at nside {ns_c} ({dt:.2f} s); ring-based C/OpenMP restatement, not libsharp"}
        out["w_production"]["gpu_same_size"] = {"ms": t_g, "value": 1e3 / t_g, "speedup_vs_cpu": dt * 1e3 / t_g,
                                                "max_rel_diff_vs_cpu": float(np.max(np.abs(dac.cpu().numpy() - a_cpu)) / np.max(np.abs(a_cpu)))}
    return out


# ------------------------------------------------------------------------------------------
# BASELINE configs[4]: MCM scaling sweep lmax 767 -> 12287 at N GPUs beside the CPU arm on the box's host cores
#   python bench.py --sweep [--gpus N under torchrun]      one JSON line per (lmax, kind)
# ------------------------------------------------------------------------------------------
def run_sweep(args):
    import torch
    import torch.distributed as dist

    import powerspectra_jl_b200 as ps
    from powerspectra_jl_b200 import device as dev
    from powerspectra_jl_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")
    L, DP = ps.lib(), ps._lib.DP
    po = cores = None
    if rank == 0 and not args.no_cpu:
        po, cores = cpu_oracle()
    reps = 5
    for lmax in (767, 1535, 3071, 6143, 12287):
        N = lmax + 1
        Vh = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
        V = torch.tensor(Vh, device="cuda")
        X = torch.empty((N, N), dtype=torch.float64, device="cuda")
        X2 = torch.empty_like(X)
        edges = dev.band_edges(0, lmax, world)
        lo, hi = edges[rank], edges[rank + 1]
        for kind, name, fam in ((0, "TT", 1), (4, "EE_BB (M++, M--)", 2)):
            def step():
                dev.mcm_slab(kind, 0, lmax, V, X, X2 if kind == 4 else None, lo, hi)
                for Xo in ((X, X2) if kind == 4 else (X,)):
                    dev.gather_bands(Xo, edges, 0, rank, world)
                    if rank == 0:
                        dev.finish(Xo, 0, lmax, True)
            for _ in range(3):
                step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.barrier(group=cpu_group)
            if rank == 0:
                terms = fam * t_fam(lmax)
                H = [torch.empty((N, N), dtype=torch.float64).pin_memory().numpy() for _ in range(2 if kind == 4 else 1)]

                def e2e():
                    ps._lib.check(L.psb200_mcm(kind, 0, lmax, Vh.ctypes.data_as(DP), Vh.size, H[0].ctypes.data_as(DP), N,
                                               H[1].ctypes.data_as(DP) if kind == 4 else None, world))
                e2e()
                t0 = time.perf_counter()
                for _ in range(3):
                    e2e()
                t_e = (time.perf_counter() - t0) * 1e3 / 3
                rec = {"sweep": "BASELINE configs[4]", "lmax": lmax, "kind": name, "n_gpus": world, "ms_resident": float(ms),
                       "terms_per_s": terms / (float(ms) * 1e-3), "e2e_ms": t_e, "e2e_terms_per_s": terms / (t_e * 1e-3),
                       "ref_terms": terms}
                if po is not None:
                    rstep = 1 if lmax <= 1535 else (8 if lmax <= 3071 else (32 if lmax <= 6143 else 128))
                    t0 = time.perf_counter()
                    tn = 0
                    for k in ((2, 3) if kind == 4 else (kind,)):
                        _, t = po.mcm(k, 0, lmax, Vh, row0=rstep // 2, rstep=rstep, return_terms=True)
                        tn += t
                    dt = time.perf_counter() - t0
                    rec["cpu_baseline"] = {"value": tn / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": f"every {rstep}th l1 row from {rstep // 2} ({tn:.3e} terms, {dt:.2f} s)"}
                    rec["speedup_vs_cpu"] = rec["terms_per_s"] / (tn / dt)
                    rec["speedup_vs_cpu_e2e"] = rec["e2e_terms_per_s"] / (tn / dt)
                print(json.dumps(rec), flush=True)
                del H
            if world > 1:
                dist.barrier(group=cpu_group)
        del X, X2
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lmax", type=int, default=6143)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra block (master, TE, QuickPol, lmax 12287)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: MCM lmax sweep 767 -> 12287, one JSON line per point")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args, args.lmax)
    else:
        run_gpu(args, args.lmax)


if __name__ == "__main__":
    main()
