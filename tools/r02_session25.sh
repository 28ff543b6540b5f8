#!/bin/bash
# Round 2, call 25: the multiprecision known-answer tests on the GPU (MCM entries at lmax 6143 / 12287, covariance entries
# at 6143) and a short bench line with the fail-safe pageable / interleaved legs.  The library is the one of call 24.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_highl_golden.py tests/test_gpu_parity.py -m gpu -q -s -k "entries or result_in_library" > gpurun_out/r02_s25_highl_tests.log 2>&1; echo "tests rc=$?"; grep "50-digit" gpurun_out/r02_s25_highl_tests.log; tail -3 gpurun_out/r02_s25_highl_tests.log
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_s25_bench.json 2> gpurun_out/r02_s25_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_s25_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_s25_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"].get("pageable_outputs"), d["e2e"].get("host_arrays"))
P
