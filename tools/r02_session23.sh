#!/bin/bash
# Round 2, call 23: staged delivery into pageable result arrays + psb200_host_alloc -- the GPU tests that go through the
# host calls, the pageable probe, a short bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "host-call tests"
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_c_abi.py -m gpu -q -x -k "lmax767 or edge_shapes or result_in_library or fused_master or short_and_rough or blocks_lmax255 or c_program or error_codes or low_rows or identities_full" > gpurun_out/r02_s23_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_s23_tests.log
step "pageable probe"
timeout 120 python tools/pageable_probe.py 6143 > gpurun_out/r02_s23_pageable_probe.jsonl 2> gpurun_out/r02_s23_pageable_probe.err; echo "probe rc=$?"; cat gpurun_out/r02_s23_pageable_probe.jsonl; tail -3 gpurun_out/r02_s23_pageable_probe.err
step "bench (short)"
timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_s23_bench.json 2> gpurun_out/r02_s23_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s23_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_s23_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"].get("pageable_outputs"), d["e2e"].get("host_arrays"), d["multi_gpu_check"]["bitwise_equal"])
P
step "done"
