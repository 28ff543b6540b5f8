#!/bin/bash
# Round 2, call 11: final single-GPU validation after the host-path changes: whole GPU suite, bench line, lmax sweep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_s11_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s11_gpu_tests.log
step "bench"
timeout 600 python bench.py > gpurun_out/r02_s11_bench.json 2> gpurun_out/r02_s11_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s11_bench.err; tail -1 gpurun_out/r02_s11_bench.json | cut -c1-400
step "sweep"
timeout 600 python bench.py --sweep > gpurun_out/r02_s11_sweep_n1.jsonl 2> gpurun_out/r02_s11_sweep.err; echo "sweep rc=$?"; cut -c1-330 gpurun_out/r02_s11_sweep_n1.jsonl; tail -2 gpurun_out/r02_s11_sweep.err
step "trace 1 GPU"
timeout 120 python tools/e2e_probe.py 6143 1 0 > gpurun_out/r02_trace_n1_tt.log 2>&1; grep -E "band:|delivered|total" gpurun_out/r02_trace_n1_tt.log | tail -8
timeout 120 python tools/e2e_probe.py 6143 1 4 > gpurun_out/r02_trace_n1_eebb.log 2>&1; grep -E "band:|delivered|total" gpurun_out/r02_trace_n1_eebb.log | tail -8
step "done"
