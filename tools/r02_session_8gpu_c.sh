#!/bin/bash
# Round 2, 8-GPU call C (final tree): the bench exactly as the driver launches it at N = 8 (extras included).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "N=8, driver flags"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n8_final.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n8_final.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"], d["config"]["pair_kernel_ms_per_rank"])
print(json.dumps(d.get("extra", {}).get("quickpol"))[:300])
print(json.dumps(d.get("extra", {}).get("w_production"))[:300])
P
step "done"
