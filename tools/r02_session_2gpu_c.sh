#!/bin/bash
# Round 2, 2-GPU call C (final tree): the multi-GPU tests (host call across GPUs, QuickPol with the flattened mapping
# across GPUs, device-side solves with peer stores) and the bench exactly as the driver launches it at N = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "multi-GPU tests"
timeout 600 python -m pytest tests -m gpu -q -k "across or two_gpus or ngpus or multi" > gpurun_out/r02_2gpu_c_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_2gpu_c_tests.log
step "bench N=2, default flags"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n2_final.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n2_final.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"], d["config"]["pair_kernel_ms_per_rank"])
print(json.dumps(d.get("extra", {}).get("quickpol"))[:400])
print(json.dumps(d.get("extra", {}).get("w_production"))[:600])
P
step "done"
