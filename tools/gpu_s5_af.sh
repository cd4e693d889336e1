#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q 2>&1 | tail -2
timeout 300 python tools/config_bench.py neg_shared max-margin 16384 256 1 30 norm 2>&1 | grep -v Warn
timeout 300 python tools/config_bench.py neg_shared skip-gram 16384 256 1 30 2>&1 | grep -v Warn
timeout 300 python tools/config_bench.py neg_shared skip-gram 2048 256 9 200 2>&1 | grep -v Warn
