#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 37 2000 2>&1 | grep -v Warn; done
timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 1 2000 2>&1 | grep -v Warn
