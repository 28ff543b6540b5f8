#!/bin/bash
# Round 2, call 9: final single-GPU evidence of the default build: whole GPU suite, smoke, bench (+extras), reference arm,
# the ncu launch list of the bench command and one ncu --set full capture of every headline pair kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_s9_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s9_gpu_tests.log
step "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s9_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s9_smoke.log
step "bench"
timeout 600 python bench.py > gpurun_out/r02_s9_bench.json 2> gpurun_out/r02_s9_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s9_bench.err; tail -1 gpurun_out/r02_s9_bench.json | cut -c1-500
step "reference arm"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_s9_bench_ref.json 2> gpurun_out/r02_s9_bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/r02_s9_bench_ref.json | cut -c1-300
step "launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s9_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_s9_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
step "ncu full"
PROBE_ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_v4 -f -o gpurun_out/r02_ncu_v4_final python tools/kernel_probe.py final > gpurun_out/r02_s9_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
