"""FP64 instructions per pair-step of every tuned pair kernel, counted in the SASS of the built library.

    python tools/sass_fp64.py [libpsb200.so] [--json out.json] [--listing out.txt]

For each `psb::pair_kernel_v3<JOB>` (`pair_kernel_v2` under PSB200_KERNEL=v2) the innermost loop with the most FP64-pipe instructions is the main
loop body: one group of R l3-steps of a thread's R pairs = R*R pair-steps.  DFMA / DMUL / DADD in that body
divided by R*R is the executed FP64 instruction count per pair-step that bench.py multiplies with the
executed pair-steps (psb200_job_stats) to get the live roofline fraction.  R is read from the library
(psb200_job_stats), not assumed.  The optional listing is the SASS of those loop bodies (committed under
profiles/ so the count can be checked by eye).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import subprocess
import sys

JOBS = ["M00", "M02", "Mpp", "Mmm", "Mpp_Mmm", "TTTT", "EEEE", "TTTE", "TETE", "TEEE_planck", "TEEE", "TTEE", "master"]
API = {"M00": (0, 0), "M02": (0, 1), "Mpp": (0, 2), "Mmm": (0, 3), "Mpp_Mmm": (0, 4), "TTTT": (1, 0), "EEEE": (1, 1),
       "TTTE": (1, 2), "TETE": (1, 3), "TEEE_planck": (1, 4), "TEEE": (1, 5), "TTEE": (1, 6), "master": (2, 0)}
FP64 = ("DFMA", "DMUL", "DADD")
# which tuned kernel the library runs by default (PSB200_KERNEL=v2 selects the recurrence kernel: count that one then)
KERNEL = "v2" if os.environ.get("PSB200_KERNEL") == "v2" else "v3"
INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")


def cuobjdump():
    for c in ("cuobjdump", "/usr/local/cuda/bin/cuobjdump"):
        try:
            subprocess.run([c, "--version"], capture_output=True, check=True)
            return c
        except (OSError, subprocess.CalledProcessError):
            continue
    return None


def functions(lib):
    exe = cuobjdump()
    if exe is None:
        raise RuntimeError("cuobjdump not found")
    txt = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = INS.match(ln)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def main_loop(ins):
    """(start, end, counts) of the innermost loop with the most FP64 instructions."""
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    best = None
    for lo, hi in loops:
        if any((a >= lo and b <= hi) and (a, b) != (lo, hi) for a, b in loops):
            continue                                     # not innermost
        body = [t for a, t in ins if lo <= a <= hi]
        cnt = {k: sum(1 for t in body if re.search(r"(^|\s)" + k + r"\b", t.split("(")[0].split(",")[0] + " ")) for k in FP64}
        n = sum(cnt.values())
        if best is None or n > best[2]:
            best = (lo, hi, n, cnt, len(body), body)
    return best


def opcode_count(body, op):
    n = 0
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        if t.split()[0].split(".")[0] == op:
            n += 1
    return n


def analyse(lib):
    L = C.CDLL(lib)
    L.psb200_job_stats.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_longlong)]
    fns = functions(lib)
    res, listing = {}, []
    for k, name in enumerate(JOBS):
        key = [f for f in fns if f"pair_kernel_{KERNEL}ILi{k}E" in f]
        if not key:
            continue
        ins = fns[key[0]]
        st = (C.c_longlong * 8)()
        api, code = API[name]
        L.psb200_job_stats(api, code, 64, 65, 0, 65, st)
        R, NR, DS = int(st[3]), int(st[4]), int(st[5])
        lo, hi, n, cnt, nbody, body = main_loop(ins)
        cnt = {op: opcode_count(body, op) for op in FP64}
        n = sum(cnt.values())
        lds = sum(opcode_count(body, op) for op in ("LDS", "LDSM"))
        res[name] = {"R": R, "rows_per_warp": NR, "l3_stride": DS, "loop": [hex(lo), hex(hi)], "instr_in_group": nbody,
                     "fp64_in_group": n, **{op.lower(): c for op, c in cnt.items()}, "lds_in_group": lds,
                     "fp64_per_pair_step": n / (R * R), "flops_per_pair_step": (2 * cnt["DFMA"] + cnt["DMUL"] + cnt["DADD"]) / (R * R),
                     "total_instr_in_kernel": len(ins)}
        listing.append(f"==== {name}: {key[0]}  main loop {hex(lo)}..{hex(hi)}  R={R} NR={NR} DS={DS}  "
                       f"{n} FP64 / {nbody} instr per group of {R * R} pair-steps")
        listing += [f"  /*{a:04x}*/ {t}" for a, t in ins if lo <= a <= hi]
    return res, listing


if __name__ == "__main__":
    args, skip = [], False
    for a in sys.argv[1:]:                      # positional = the library; --json / --listing take a value
        if skip:
            skip = False
        elif a in ("--json", "--listing"):
            skip = True
        elif not a.startswith("--"):
            args.append(a)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = args[0] if args else os.path.join(here, "powerspectra.jl_b200", "libpsb200.so")
    res, listing = analyse(lib)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(res, f, indent=1)
    if "--listing" in sys.argv:
        with open(sys.argv[sys.argv.index("--listing") + 1], "w") as f:
            f.write("\n".join(listing) + "\n")
    for k, v in res.items():
        print(f"{k:12s} R={v['R']} NR={v['rows_per_warp']} DS={v['l3_stride']}  FP64/pair-step {v['fp64_per_pair_step']:.3f} "
              f"(DFMA {v['dfma']} DMUL {v['dmul']} DADD {v['dadd']} per {v['R']**2}), LDS {v['lds_in_group']}, "
              f"{v['instr_in_group']} instr per group")
