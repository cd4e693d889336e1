#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q -k "fused or self_gather" 2>&1 | tail -2
( for sg in 1 0 1 0; do echo "NNCF_SELF_GATHER=$sg"; NNCF_SELF_GATHER=$sg timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 37 2000; done
  for sg in 1 0; do echo "NNCF_SELF_GATHER=$sg"; NNCF_SELF_GATHER=$sg timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 1 2000; done ) 2>&1 | grep -v Warning | tee gpurun_out/s5u_configs.txt
