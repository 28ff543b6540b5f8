#!/bin/bash
# Round 2, call 4: device-side decoupling tests (psb200_mcm_solve / master_solve / decouple_covmat), whole GPU suite
# on the x-column build, bench line with extras.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "solve tests"
timeout 300 python -m pytest tests/test_gpu_solve.py -m gpu -x -q > gpurun_out/r02_s4_solve_tests.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02_s4_solve_tests.log
step "gpu suite"
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_solve.py > gpurun_out/r02_s4_gpu_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s4_gpu_tests.log
step "bench"
timeout 500 python bench.py > gpurun_out/r02_s4_bench.json 2> gpurun_out/r02_s4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s4_bench.err; tail -1 gpurun_out/r02_s4_bench.json | cut -c1-1200
step "done"
