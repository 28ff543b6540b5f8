#!/bin/bash
# Round 2, call 20: analysis Legendre kernel with two 8-step reductions per pass (96 registers, 5 blocks per SM) vs one of 16.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "sht tests (V = 8 default)"
timeout 900 python -m pytest tests/test_sht.py -m gpu -q > gpurun_out/r02_s20_sht_tests.log 2>&1; echo "sht_tests rc=$?"; tail -4 gpurun_out/r02_s20_sht_tests.log
step "sht probe"
: > gpurun_out/r02_s20_sht_probe.jsonl
for V in 8 16; do
  PSB200_SHT_V=$V timeout 300 python tools/sht_probe.py 2048 >> gpurun_out/r02_s20_sht_probe.jsonl 2>> gpurun_out/r02_s20_sht_probe.err
  PSB200_SHT_V=$V PSB200_SHT_R=8 timeout 300 python tools/sht_probe.py 2048 >> gpurun_out/r02_s20_sht_probe.jsonl 2>> gpurun_out/r02_s20_sht_probe.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_s20_sht_probe.jsonl'):
    d = json.loads(ln)
    print(d['nside'], 'R', d['stats']['R'], 'syn %.2f ana %.2f m2a %.2f ms' % (d['ms_alm2map'], d['ms_analysis'], d['ms_map2alm_niter3']), 'frac syn %.3f ana %.3f' % (d['frac_synthesis_pass'], d['frac_analysis_pass']))
PY
tail -3 gpurun_out/r02_s20_sht_probe.err
step "done"
