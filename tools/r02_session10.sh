#!/bin/bash
# Round 2, call 10: host-level calls with sub-bands alternating between two streams: whole GPU suite, e2e A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_s10_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s10_gpu_tests.log
step "bench two streams"
timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s10_bench_2s.json 2> gpurun_out/r02_s10_bench_2s.err; echo "rc=$?"; tail -2 gpurun_out/r02_s10_bench_2s.err
step "bench one stream"
PSB200_ONE_STREAM=1 timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s10_bench_1s.json 2> gpurun_out/r02_s10_bench_1s.err; echo "rc=$?"; tail -2 gpurun_out/r02_s10_bench_1s.err
for n in 4 16; do
step "bench two streams nsub=$n"
PSB200_NSUB=$n timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s10_bench_2s_nsub$n.json 2> /dev/null; echo "rc=$?"
done
python - <<'P'
import json
for v in ("2s","1s","2s_nsub4","2s_nsub16"):
    try:
        d=json.loads(open(f"gpurun_out/r02_s10_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"]["bitwise_equal"])
    except Exception as e: print(v, "failed", e)
P
step "done"
