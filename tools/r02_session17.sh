#!/bin/bash
# Round 2, call 17: short-ring fold of the ring synthesis; QuickPol flattened thread mapping A/B; whole GPU suite and the
# bench line (with the new w_production extra) on the current tree.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "sht tests"
timeout 600 python -m pytest tests/test_sht.py -m gpu -q > gpurun_out/r02_s17_sht_tests.log 2>&1; echo "sht_tests rc=$?"; tail -4 gpurun_out/r02_s17_sht_tests.log
step "sht probe"
timeout 300 python tools/sht_probe.py 2048 > gpurun_out/r02_s17_sht_probe.jsonl 2> gpurun_out/r02_s17_sht_probe.err
timeout 300 python tools/sht_probe.py 1024 >> gpurun_out/r02_s17_sht_probe.jsonl 2>> gpurun_out/r02_s17_sht_probe.err
cut -c1-330 gpurun_out/r02_s17_sht_probe.jsonl; tail -3 gpurun_out/r02_s17_sht_probe.err
step "launch list nside 2048"
PROBE_ONCE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sht_ -c 60 --csv --log-file gpurun_out/r02_s17_sht_launches.csv python tools/sht_probe.py 2048 > gpurun_out/r02_s17_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02_s17_sht_launches.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[:11]:
    print(r[4][:50], r[-1])
PY
step "quickpol flat A/B"
timeout 200 python tests/tools/quickpol_probe.py 6143 128 gpurun_out/r02_s17_qp_default.json > /dev/null 2> gpurun_out/r02_s17_qp.err; echo "rc=$?"
PSB200_LIB=$PWD/tools/_build/libpsb200_qpflat.so timeout 200 python tests/tools/quickpol_probe.py 6143 128 gpurun_out/r02_s17_qp_flat.json > /dev/null 2>> gpurun_out/r02_s17_qp.err; echo "rc=$?"
python - <<'PY'
import json
for t in ("default", "flat"):
    try:
        d = json.load(open(f"gpurun_out/r02_s17_qp_{t}.json"))
        print(t, {k: v for k, v in d.items() if "ms" in k or "err" in k or "parity" in k})
    except Exception as e:
        print(t, "failed", e)
PY
PSB200_LIB=$PWD/tools/_build/libpsb200_qpflat.so timeout 300 python -m pytest tests/test_quickpol.py -m gpu -q > gpurun_out/r02_s17_qp_flat_tests.log 2>&1; echo "flat tests rc=$?"; tail -3 gpurun_out/r02_s17_qp_flat_tests.log
step "gpu suite"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_s17_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -5 gpurun_out/r02_s17_gpu_tests.log
step "bench"
timeout 900 python bench.py > gpurun_out/r02_s17_bench.json 2> gpurun_out/r02_s17_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s17_bench.err; tail -1 gpurun_out/r02_s17_bench.json | cut -c1-400
step "done"
