#!/bin/bash
# round 2, call V (1 GPU): teams score kernel (4 S' buffers of 32 columns) vs the 2 x 64 build: parity tests, then step times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -5
for lib in "" build/ablate/libnncf_noteams.so; do
  echo "=== lib ${lib:-teams (default)}"
  for cfg in "neg_shared skip-gram 512 128 37 2000 ureg" "neg_shared skip-gram 512 128 1 3000 ureg" "neg_shared skip-gram 512 128 37 1000 ureg adam" "neg_shared log-loss 512 128 37 1000 ureg norm" "group_neg_shared log-loss 512 128 37 1000 ureg norm" "neg_shared skip-gram 4096 128 5 500 ureg" "neg_shared skip-gram 512 64 37 1000 ureg"; do
    NNCF_LIB_PATH=$lib ZIPF=10,10 timeout 120 python tools/config_bench.py $cfg 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/r02v_teams.txt
