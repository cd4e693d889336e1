#!/bin/bash
mkdir -p gpurun_out
for pf in 1 0 1 0; do echo "NNCF_PREFETCH=$pf"; NNCF_PREFETCH=$pf timeout 300 python tools/config_bench.py neg_shared skip-gram 512 128 37 500 adam 2>&1 | grep -v Warn; done
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q -k "adam" 2>&1 | tail -2
