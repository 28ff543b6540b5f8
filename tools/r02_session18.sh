#!/bin/bash
# Round 2, call 18: un-normalised Legendre recurrence (4 instead of 5 FP64 instructions per step), two-stage alm2cl:
# tests, probe at nside 2048 / 1024, launch list, ncu --set full at nside 1024.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "sht tests"
timeout 900 python -m pytest tests/test_sht.py -m gpu -q > gpurun_out/r02_s18_sht_tests.log 2>&1; echo "sht_tests rc=$?"; tail -6 gpurun_out/r02_s18_sht_tests.log
step "sht probe"
: > gpurun_out/r02_s18_sht_probe.jsonl
for R in 4 8; do
  PSB200_SHT_R=$R timeout 300 python tools/sht_probe.py 2048 >> gpurun_out/r02_s18_sht_probe.jsonl 2>> gpurun_out/r02_s18_sht_probe.err
  tail -1 gpurun_out/r02_s18_sht_probe.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('R',d['stats']['R'],'syn %.2f ana %.2f m2a %.2f ms'%(d['ms_alm2map'],d['ms_analysis'],d['ms_map2alm_niter3']),'frac syn %.3f ana %.3f'%(d['frac_synthesis_pass'],d['frac_analysis_pass']))"
done
timeout 200 python tools/sht_probe.py 1024 >> gpurun_out/r02_s18_sht_probe.jsonl 2>> gpurun_out/r02_s18_sht_probe.err
tail -3 gpurun_out/r02_s18_sht_probe.err
step "launch list nside 2048"
PROBE_ONCE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sht_ -c 60 --csv --log-file gpurun_out/r02_s18_sht_launches.csv python tools/sht_probe.py 2048 > gpurun_out/r02_s18_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02_s18_sht_launches.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[:11]:
    print(r[4][:50], r[-1])
PY
step "ncu full, nside 1024"
PROBE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sht_ -c 11 -f -o gpurun_out/r02_ncu_sht3 python tools/sht_probe.py 1024 > gpurun_out/r02_s18_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_s18_ncu.log
step "done"
