#!/bin/bash
# Round 2, call 13: first run of the spin-0 HEALPix transforms (psb200_sht.cuh) on the device: parity tests, timing
# probe with R = 2 / 4 / 8 ring pairs per lane, then the whole GPU suite and a bench line on the re-created tree.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "sht tests"
timeout 600 python -m pytest tests/test_sht.py -m gpu -q -x > gpurun_out/r02_s13_sht_tests.log 2>&1; echo "sht_tests rc=$?"; tail -25 gpurun_out/r02_s13_sht_tests.log
step "sht probe"
for R in 4 2 8; do
  PSB200_SHT_R=$R timeout 200 python tools/sht_probe.py 1024 >> gpurun_out/r02_s13_sht_probe.jsonl 2>> gpurun_out/r02_s13_sht_probe.err
  PSB200_SHT_R=$R timeout 300 python tools/sht_probe.py 2048 >> gpurun_out/r02_s13_sht_probe.jsonl 2>> gpurun_out/r02_s13_sht_probe.err
done
cat gpurun_out/r02_s13_sht_probe.jsonl; tail -5 gpurun_out/r02_s13_sht_probe.err
step "gpu suite"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_s13_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -8 gpurun_out/r02_s13_gpu_tests.log
step "done"
