#!/bin/bash
# Short 1-GPU call after the QuickPol loop restructuring: parity (both instantiations), timing, ncu of the tab kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "quickpol gpu tests"
timeout 100 python -m pytest tests/test_quickpol.py -m gpu -x -q > gpurun_out/qp_tests2.log 2>&1; echo "qp_tests rc=$?"; tail -3 gpurun_out/qp_tests2.log
step "quickpol probe"
timeout 60 python tests/tools/quickpol_probe.py 6143 128 gpurun_out/quickpol_probe2.json > gpurun_out/qp_probe2.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/qp_probe2.log | cut -c1-1200
step "ncu quickpol tab"
QP_PROBE_VARIANTS=tab QP_PROBE_FAST=1 timeout 60 ncu --set full --clock-control none --import-source on -k regex:quickpol_kernel \
  --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r01_quickpol_tab2 python tests/tools/quickpol_probe.py 6143 128 > gpurun_out/ncu_qp_tab2.log 2>&1; echo "ncu tab rc=$?"
step "smoke"
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
step "done"
