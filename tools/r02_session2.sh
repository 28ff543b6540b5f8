#!/bin/bash
# Round 2, call 2: closed-form kernel (v3) -- GPU suite, bench line, and the v2 (recurrence) kernel timed beside it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu parity tests (v3 default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_s2_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -15 gpurun_out/r02_s2_gpu_tests.log
step "bench v3 (no cpu, no extra)"
timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s2_bench_v3.json 2> gpurun_out/r02_s2_bench_v3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s2_bench_v3.err
step "bench v2 (no cpu, no extra)"
PSB200_KERNEL=v2 timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s2_bench_v2.json 2> gpurun_out/r02_s2_bench_v2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s2_bench_v2.err
python - <<'P'
import json
for v in ("v3","v2"):
    try:
        d=json.loads(open(f"gpurun_out/r02_s2_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), {k:round(x["ms"],2) for k,x in d["roofline"]["all_kernels"].items()}, {k:round(x["frac"],3) for k,x in d["roofline"]["all_kernels"].items()}, d["multi_gpu_check"]["bitwise_equal"])
    except Exception as e: print(v, "failed", e)
P
step "done"
