#!/bin/bash
# Round 2, call 5: ring-staged closed-form kernel (v4): parity, A/B against v3, tiling variants, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu parity tests (v4 default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_s5_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -15 gpurun_out/r02_s5_gpu_tests.log
: > gpurun_out/r02_s5_probe.jsonl
step "probe v3"
PSB200_KERNEL=v3 PSB200_LIB=$PWD/tools/_build/libpsb200_base.so timeout 120 python tools/kernel_probe.py v3base >> gpurun_out/r02_s5_probe.jsonl 2> gpurun_out/r02_s5_probe_v3.err || echo "probe v3 failed"
tail -1 gpurun_out/r02_s5_probe.jsonl | cut -c1-300
for v in base noxcol rh8 rall8 nr1 nr4 minb8 rh4; do
  step "probe $v"
  PSB200_LIB=$PWD/tools/_build/libpsb200_$v.so timeout 120 python tools/kernel_probe.py $v >> gpurun_out/r02_s5_probe.jsonl 2> gpurun_out/r02_s5_probe_$v.err || echo "probe $v failed"
  tail -1 gpurun_out/r02_s5_probe.jsonl | cut -c1-300
done
step "ncu base"
PROBE_ONCE=1 PSB200_LIB=$PWD/tools/_build/libpsb200_base.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_v4 -f -o gpurun_out/r02_ncu_v4_base python tools/kernel_probe.py base > gpurun_out/r02_s5_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
