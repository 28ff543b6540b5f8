#!/bin/bash
# One gpurun call (1 GPU): QuickPol parity on the device (both kernel instantiations), its timing + ncu evidence,
# the DFMA/DMMA concurrency probe, smoke(), the whole GPU suite, a bench line.  Every step has its own timeout and
# writes under gpurun_out/; later steps run even if an earlier one fails.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "quickpol gpu tests"
timeout 150 python -m pytest tests/test_quickpol.py -m gpu -x -q > gpurun_out/qp_tests.log 2>&1; echo "qp_tests rc=$?"; tail -3 gpurun_out/qp_tests.log
step "dmma probe"
[ -x tools/_build/dmma_probe ] || { mkdir -p tools/_build; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/dmma_probe tools/dmma_probe.cu; }
timeout 30 tools/_build/dmma_probe > gpurun_out/dmma_probe.json 2>&1; echo "dmma rc=$?"; cat gpurun_out/dmma_probe.json
step "quickpol probe"
timeout 100 python tests/tools/quickpol_probe.py 6143 128 gpurun_out/quickpol_probe.json > gpurun_out/qp_probe.log 2>&1; echo "probe rc=$?"; tail -2 gpurun_out/qp_probe.log | cut -c1-1500
step "smoke"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
step "ncu quickpol tab"
QP_PROBE_VARIANTS=tab QP_PROBE_FAST=1 timeout 80 ncu --set full --clock-control none --import-source on -k regex:quickpol_kernel \
  --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r01_quickpol_tab python tests/tools/quickpol_probe.py 6143 128 > gpurun_out/ncu_qp_tab.log 2>&1; echo "ncu tab rc=$?"
step "full gpu suite (quickpol file already run)"
timeout 240 python -m pytest tests -m gpu -x -q --ignore=tests/test_quickpol.py > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
step "ncu quickpol simple"
QP_PROBE_VARIANTS=simple QP_PROBE_FAST=1 timeout 80 ncu --set full --clock-control none --import-source on -k regex:quickpol_kernel \
  --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r01_quickpol_simple python tests/tools/quickpol_probe.py 6143 128 > gpurun_out/ncu_qp_simple.log 2>&1; echo "ncu simple rc=$?"
step "bench"
timeout 150 python bench.py > gpurun_out/bench_r01_s2_n1.json 2> gpurun_out/bench_r01_s2_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01_s2_n1.json | cut -c1-600
step "done"
