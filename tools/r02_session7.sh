#!/bin/bash
# Round 2, call 7: v4 variants round 3 (new default tiling, elements per lane, x columns), parity + ncu of the default build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
: > gpurun_out/r02_s7_probe.jsonl
for v in cfgA epl2 noxcol both2 light4 epl2rb8 mid6; do
  step "probe $v"
  PSB200_LIB=$PWD/tools/_build/libpsb200_$v.so timeout 120 python tools/kernel_probe.py $v >> gpurun_out/r02_s7_probe.jsonl 2> gpurun_out/r02_s7_probe_$v.err || echo "probe $v failed"
  tail -1 gpurun_out/r02_s7_probe.jsonl | cut -c1-260
done
step "gpu parity tests (default build = cfgA)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_s7_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -5 gpurun_out/r02_s7_gpu_tests.log
step "parity epl2"
PSB200_LIB=$PWD/tools/_build/libpsb200_epl2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not lmax12287 and not 6143" > gpurun_out/r02_s7_gpu_tests_epl2.log 2>&1; echo "epl2 tests rc=$?"; tail -3 gpurun_out/r02_s7_gpu_tests_epl2.log
step "ncu cfgA"
PROBE_ONCE=1 PSB200_LIB=$PWD/tools/_build/libpsb200_cfgA.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_v4 -f -o gpurun_out/r02_ncu_v4_cfgA python tools/kernel_probe.py cfgA > gpurun_out/r02_s7_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
