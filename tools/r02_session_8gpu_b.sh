#!/bin/bash
# Round 2, 8-GPU call B: folded split (a low and a high band per rank, one tile list) against one contiguous band per rank,
# then the bench exactly as the driver launches it (extras included).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
show() { python - <<P
import json
d=json.loads(open("$1").read().strip().splitlines()[-1])
print("$1: ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "driver", d["e2e"]["per_rank_driver"] and round(d["e2e"]["per_rank_driver"]["ms_per_step"],2), d["multi_gpu_check"]["bitwise_equal"], d["config"]["pair_kernel_ms_per_rank"])
P
}
step "N=8 contiguous"
PSB200_BENCH_SPLIT=contiguous timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n8_contiguous.json 2> gpurun_out/r02_bench_n8_contiguous.err; echo "rc=$?"; show gpurun_out/r02_bench_n8_contiguous.json
step "N=8 folded, driver flags"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8_folded.json 2> gpurun_out/r02_bench_n8_folded.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n8_folded.err | cut -c1-300; show gpurun_out/r02_bench_n8_folded.json
step "N=4 folded"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n4_folded.json 2> gpurun_out/r02_bench_n4_folded.err; echo "rc=$?"; show gpurun_out/r02_bench_n4_folded.json
step "done"
