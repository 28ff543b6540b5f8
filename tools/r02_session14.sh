#!/bin/bash
# Round 2, call 14: where the time of the HEALPix transforms goes: launch list at nside 2048, ncu --set full of the two
# Legendre kernels and the two ring kernels at nside 1024; the fixed full-size property test.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "full-size property test"
timeout 600 python -m pytest tests/test_sht.py -m gpu -q -k full_size > gpurun_out/r02_s14_sht_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_s14_sht_tests.log
step "launch list nside 2048"
PROBE_ONCE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s14_sht_launches.csv python tools/sht_probe.py 2048 > gpurun_out/r02_s14_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_s14_sht_launches.csv')) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4][:60]
    agg.setdefault(k, []).append(float(r[-1].replace(',', '')))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} total={sum(v)/1e6:9.3f} ms  max={max(v)/1e6:8.3f} ms")
PY
step "ncu full, nside 1024"
PROBE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sht_ -c 14 -f -o gpurun_out/r02_ncu_sht python tools/sht_probe.py 1024 > gpurun_out/r02_s14_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_s14_ncu.log
step "done"
