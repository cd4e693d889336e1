#!/bin/bash
# round 2, call W (1 GPU): G' exchange (two-sided score kernel): parity, then step times with the exchange on / off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k "g_exchange" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -5
for gx in 1 0; do
  echo "=== NNCF_GX=$gx"
  for cfg in "neg_shared skip-gram 512 128 37 2000 ureg" "neg_shared skip-gram 512 128 1 3000 ureg" "neg_shared skip-gram 512 128 37 1000 ureg adam" "neg_shared mse 512 128 37 1000 ureg" "neg_shared skip-gram 4096 128 5 500 ureg" "neg_shared skip-gram 512 64 37 1000 ureg" "neg_shared skip-gram 512 128 74 1000 ureg"; do
    NNCF_GX=$gx ZIPF=10,10 timeout 120 python tools/config_bench.py $cfg 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/r02w_gx.txt
