#!/bin/bash
# Round 2, call 12: tile lists over several bands (int4 entries): whole GPU suite + a quick bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_s12_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -8 gpurun_out/r02_s12_gpu_tests.log
step "bench quick"
timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_s12_bench.json 2> gpurun_out/r02_s12_bench.err; echo "rc=$?"; tail -2 gpurun_out/r02_s12_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_s12_bench.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['multi_gpu_check']['bitwise_equal'])"
step "done"
