#!/bin/bash
# Round 2, 2-GPU call B: the bench exactly as the driver launches it at N = 2 (extras included), and the sweep at N = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "bench N=2, default flags"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2_full.json 2> gpurun_out/r02_bench_n2_full.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n2_full.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n2_full.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"]["bitwise_equal"], d["config"]["pair_kernel_ms_per_rank"], d["config"]["band_edges"])
print(json.dumps(d.get("extra"))[:1500])
P
step "reference arm under torchrun"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; echo "ref rc=$?"; tail -1 gpurun_out/r02_bench_ref_n2.json | cut -c1-400
step "done"
