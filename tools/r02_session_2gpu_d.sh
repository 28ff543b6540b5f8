#!/bin/bash
# Round 2, 2-GPU call D: psb200_map2alm_many across two devices (bit-identical to the one-at-a-time path); the two tests
# whose expectations were wrong in call 22 (Jacobi iterations do not keep a_00 of a constant map exact at nside 2).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sht.py tests/test_c_abi.py -m gpu -q -k "batched or argument_errors or c_program" > gpurun_out/r02_2gpu_d_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_2gpu_d_tests.log
