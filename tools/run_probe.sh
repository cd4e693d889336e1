#!/bin/bash
# runs the UMMA descriptor probe with a few descriptor variants; each in its own process, bounded by timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.log 2>&1
for args in "128 16384 1024 16 1024" "64 16384 1024 16 1024" "256 16384 1024 16 1024" \
            "128 1024 16384 16 1024" "128 16384 1024 0 1024" "128 16384 128 16 1024" "128 128 1024 16 1024"; do
  timeout 30 ./build/umma_probe $args >> gpurun_out/probe.log 2>&1
  echo "exit=$? args=$args" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
