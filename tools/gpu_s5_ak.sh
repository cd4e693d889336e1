#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q 2>&1 | tail -2
for ho in 1 0 1 0; do echo "NNCF_BLOCK_HANDOVER=$ho"; NNCF_BLOCK_HANDOVER=$ho timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 37 2000 2>&1 | grep -v Warn; done
for ho in 1 0; do echo "NNCF_BLOCK_HANDOVER=$ho"; NNCF_BLOCK_HANDOVER=$ho timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 1 2000 2>&1 | grep -v Warn; done
