#!/bin/bash
# Round 2, call 24: staged copies everywhere a host call moves a large pageable buffer (result matrices, HEALPix maps and
# alm), streaming stores in the scatter (PSB200_STAGE_NT A/B), 12-worker default: pageable probe, whole GPU suite, the
# bench line, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "pageable probe"
timeout 120 python tools/pageable_probe.py 6143 > gpurun_out/r02_s24_pageable_probe.jsonl 2> gpurun_out/r02_s24_pageable_probe.err; echo "probe rc=$?"; cat gpurun_out/r02_s24_pageable_probe.jsonl; tail -3 gpurun_out/r02_s24_pageable_probe.err
step "gpu suite"
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02_s24_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s24_gpu_tests.log
step "bench"
timeout 420 python bench.py > gpurun_out/r02_s24_bench.json 2> gpurun_out/r02_s24_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s24_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_s24_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"].get("pageable_outputs"), d["multi_gpu_check"]["bitwise_equal"])
w=d["extra"]["w_production"]; print({k: w[k] for k in ("map2alm_niter3_ms","map2alm_e2e_ms","batched")})
P
step "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s24_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s24_smoke.log
step "done"
