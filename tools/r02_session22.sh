#!/bin/bash
# Round 2, call 22: batched effective weights (psb200_map2alm_many), psb200_sht_release, plain-C client with map2alm:
# whole GPU suite, bench line, ncu --set full of the QuickPol kernel with the flattened mapping.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "gpu suite"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_s22_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s22_gpu_tests.log
step "bench"
timeout 900 python bench.py > gpurun_out/r02_s22_bench.json 2> gpurun_out/r02_s22_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s22_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_s22_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"]["bitwise_equal"])
w=d["extra"]["w_production"]; print({k: w[k] for k in ("map2alm_niter3_ms","map2alm_e2e_ms","batched")})
P
step "ncu quickpol"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:quickpol_kernel -c 1 -f -o gpurun_out/r02_ncu_quickpol_flat python tests/tools/quickpol_probe.py 6143 128 > gpurun_out/r02_s22_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
