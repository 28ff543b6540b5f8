#!/bin/bash
# Final 1-GPU validation of the shipped build: whole GPU suite, QuickPol probe + ncu evidence, launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "gpu_tests rc=$?"; tail -2 gpurun_out/gpu_tests_final.log
step "quickpol probe"
timeout 60 python tests/tools/quickpol_probe.py 6143 128 gpurun_out/quickpol_probe_final.json > gpurun_out/qp_probe_final.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/qp_probe_final.log | cut -c1-900
step "ncu quickpol tab"
QP_PROBE_VARIANTS=tab QP_PROBE_FAST=1 timeout 60 ncu --set full --clock-control none --import-source on -k regex:quickpol_kernel \
  --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r01_quickpol_final python tests/tools/quickpol_probe.py 6143 128 > gpurun_out/ncu_qp_final.log 2>&1; echo "ncu rc=$?"
step "launch list quickpol"
QP_PROBE_VARIANTS=tab QP_PROBE_FAST=1 timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_quickpol.csv python tests/tools/quickpol_probe.py 6143 128 > /dev/null 2>&1; echo "launch list rc=$?"
step "done"
