#!/bin/bash
# Round 2, call 3: tiling variants of the closed-form kernel (A/B by kernel_probe) + ncu --set full of the base build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
: > gpurun_out/r02_s3_probe.jsonl
for v in base nr4 nr1 noxcol minb8 minb12 rh8 rh4 rl6 rb8 tc64 tc256; do
  step "probe $v"
  PSB200_LIB=$PWD/tools/_build/libpsb200_$v.so timeout 120 python tools/kernel_probe.py $v >> gpurun_out/r02_s3_probe.jsonl 2> gpurun_out/r02_s3_probe_$v.err || echo "probe $v failed"
  tail -1 gpurun_out/r02_s3_probe.jsonl | cut -c1-400
done
step "ncu base"
PROBE_ONCE=1 PSB200_LIB=$PWD/tools/_build/libpsb200_base.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_v3 -f -o gpurun_out/r02_ncu_v3_base python tools/kernel_probe.py base > gpurun_out/r02_s3_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
