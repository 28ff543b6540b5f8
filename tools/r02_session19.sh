#!/bin/bash
# Round 2, call 19: synthesis Legendre kernel with deferred combine of the prefetched values and 5 blocks per SM.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "sht tests"
timeout 900 python -m pytest tests/test_sht.py -m gpu -q > gpurun_out/r02_s19_sht_tests.log 2>&1; echo "sht_tests rc=$?"; tail -4 gpurun_out/r02_s19_sht_tests.log
step "sht probe"
timeout 300 python tools/sht_probe.py 2048 > gpurun_out/r02_s19_sht_probe.jsonl 2> gpurun_out/r02_s19_sht_probe.err
timeout 300 python tools/sht_probe.py 1024 >> gpurun_out/r02_s19_sht_probe.jsonl 2>> gpurun_out/r02_s19_sht_probe.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_s19_sht_probe.jsonl'):
    d = json.loads(ln)
    print(d['nside'], 'R', d['stats']['R'], 'syn %.2f ana %.2f m2a %.2f ms' % (d['ms_alm2map'], d['ms_analysis'], d['ms_map2alm_niter3']), 'frac syn %.3f ana %.3f' % (d['frac_synthesis_pass'], d['frac_analysis_pass']))
PY
step "launch list nside 2048"
PROBE_ONCE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sht_ -c 60 --csv --log-file gpurun_out/r02_s19_sht_launches.csv python tools/sht_probe.py 2048 > gpurun_out/r02_s19_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02_s19_sht_launches.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[:11]:
    print(r[4][:50], r[-1])
PY
step "done"
