#!/bin/bash
# Round 2, call 21 (final tree): whole GPU suite, smoke, the bench line, the bench launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s21_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s21_smoke.log
step "gpu suite"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_s21_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -5 gpurun_out/r02_s21_gpu_tests.log
step "bench"
timeout 900 python bench.py > gpurun_out/r02_s21_bench.json 2> gpurun_out/r02_s21_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s21_bench.err; tail -1 gpurun_out/r02_s21_bench.json | cut -c1-300
step "bench launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s21_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_s21_launch.log 2>&1; echo "rc=$?"
step "done"
