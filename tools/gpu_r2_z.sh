#!/bin/bash
# round 2, call Z (1 GPU): the failing host-fed test with its message; host-fed vs device-fed wall clock per chunk length / loss path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k "host_fed" 2>&1 | grep -E "^E|passed|failed" | head -30
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 900 python tools/host_fed_bench.py 2>&1 | grep -v Warning | tee gpurun_out/r02z_host_fed.txt
