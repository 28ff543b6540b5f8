#!/bin/bash
# Round 2, 8-GPU call: torchrun bench at N = 8 and N = 4 (multi_gpu_check inside), phase trace of the host call with ngpus = 8.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 8 4; do
step "bench N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n$n.err | cut -c1-300
python - <<P
import json
d=json.loads(open("gpurun_out/r02_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "per-rank driver", d["e2e"]["per_rank_driver"] and round(d["e2e"]["per_rank_driver"]["ms_per_step"],2), d["multi_gpu_check"], d["config"]["pair_kernel_ms_per_rank"], d["config"]["band_edges"])
P
done
step "trace TT ngpus=8"
timeout 120 python tools/e2e_probe.py 6143 8 0 > gpurun_out/r02_trace_n8_tt.log 2>&1; grep -E "host bands|band:|delivered|total" gpurun_out/r02_trace_n8_tt.log | tail -24
step "trace EEBB ngpus=8"
timeout 120 python tools/e2e_probe.py 6143 8 4 > gpurun_out/r02_trace_n8_eebb.log 2>&1; grep -E "host bands|band:|delivered|total" gpurun_out/r02_trace_n8_eebb.log | tail -24
step "done"
