#!/bin/bash
# Round 2, 2-GPU call E: psb200_host_alloc on a box with several GPUs (and, if it has them, several NUMA nodes):
# placement, registration, host calls on all GPUs into one-node and interleaved arrays.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)") > gpurun_out/r02_2gpu_e_topo.txt 2>&1
timeout 100 python tools/numa_probe.py 6143 > gpurun_out/r02_2gpu_e_numa_probe.jsonl 2> gpurun_out/r02_2gpu_e_numa_probe.err; echo "rc=$?"
cat gpurun_out/r02_2gpu_e_numa_probe.jsonl; tail -3 gpurun_out/r02_2gpu_e_numa_probe.err; grep -i "numa\|socket" gpurun_out/r02_2gpu_e_topo.txt | head
