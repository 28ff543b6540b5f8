#!/bin/bash
# Round 2, 2-GPU call: the multi-GPU tests that a 1-GPU box skips, the torchrun bench at N = 2 with its bitwise
# multi_gpu_check, and a phase trace of the host-level call with ngpus = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader
step "multi-GPU tests"
timeout 400 python -m pytest tests -m gpu -q -k "across or two_gpus or ngpus or gpus" > gpurun_out/r02_2gpu_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_2gpu_tests.log
step "bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n2.err; tail -1 gpurun_out/r02_bench_n2.json | cut -c1-1500
step "trace TT ngpus=2"
timeout 120 python tools/e2e_probe.py 6143 2 0 > gpurun_out/r02_trace_n2_tt.log 2>&1; tail -40 gpurun_out/r02_trace_n2_tt.log
step "trace EEBB ngpus=2"
timeout 120 python tools/e2e_probe.py 6143 2 4 > gpurun_out/r02_trace_n2_eebb.log 2>&1; tail -24 gpurun_out/r02_trace_n2_eebb.log
step "1-GPU e2e with 32 sub-bands"
PSB200_NSUB=32 timeout 200 python bench.py --no-cpu --no-extra > gpurun_out/r02_2gpu_bench_n1_nsub32.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_2gpu_bench_n1_nsub32.json').read().strip().splitlines()[-1]); print('nsub32 ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
step "done"
