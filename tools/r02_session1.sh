#!/bin/bash
# Round 2, first 1-GPU call: whole GPU suite, smoke(), the default bench line, the reference arm, the ncu launch list
# of the bench command.  Every step has its own timeout and writes under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -5 gpurun_out/r02_gpu_tests.log
step "smoke"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
step "bench"
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench_n1.json | cut -c1-3000; tail -5 gpurun_out/r02_bench_n1.err
step "reference arm"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/r02_bench_ref.json | cut -c1-1200
step "launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
step "done"
