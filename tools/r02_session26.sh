#!/bin/bash
# Round 2, call 26: the bench line of the final tree (known_answers, fail-safe pageable / interleaved legs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/r02_s26_bench.json 2> gpurun_out/r02_s26_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s26_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_s26_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"].get("pageable_outputs"), d["known_answers"], d["multi_gpu_check"]["bitwise_equal"], d["roofline"]["frac"])
P
