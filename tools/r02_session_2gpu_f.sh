#!/bin/bash
# Round 2, 2-GPU call F: the staged delivery with one host thread per GPU (pageable result arrays across two devices):
# the multi-GPU tests (numpy arrays are pageable) and the pageable probe with ngpus = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_quickpol.py tests/test_gpu_solve.py tests/test_sht.py -m gpu -q -k "two_gpus or across_gpus or multi_gpu or several" > gpurun_out/r02_2gpu_f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_2gpu_f_tests.log
timeout 60 python tools/pageable_probe.py 6143 2 quick > gpurun_out/r02_2gpu_f_pageable_probe.jsonl 2> gpurun_out/r02_2gpu_f_pageable_probe.err; echo "probe rc=$?"; cat gpurun_out/r02_2gpu_f_pageable_probe.jsonl; tail -3 gpurun_out/r02_2gpu_f_pageable_probe.err
