#!/bin/bash
mkdir -p gpurun_out
echo "== tests (train step)"; timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -4
echo "== config benches: self-gather on / off"
( for sg in 1 0 1 0; do echo "NNCF_SELF_GATHER=$sg"; NNCF_SELF_GATHER=$sg timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 37 2000; done
  for sg in 1 0; do echo "NNCF_SELF_GATHER=$sg"; NNCF_SELF_GATHER=$sg timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 1 2000; done
  for sg in 1 0; do echo "NNCF_SELF_GATHER=$sg"; NNCF_SELF_GATHER=$sg timeout 200 python tools/config_bench.py neg_shared skip-gram 4096 128 4 300; done ) 2>&1 | grep -v Warning | tee gpurun_out/s5t_configs.txt
