#!/bin/bash
# session 5, call A: whole@k pipeline ablations (second generation), C5 / large-batch C3 step timings, first run of the CTA-pair kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== g2 sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 g2 2>&1 | tee gpurun_out/s5a_sweep_g2.txt | tail -20
echo "== config benches"
( timeout 200 python tools/config_bench.py neg_shared max-margin 16384 256 1 20 norm
  timeout 200 python tools/config_bench.py neg_shared skip-gram 16384 256 1 20
  timeout 200 python tools/config_bench.py neg_shared skip-gram 4096 128 5 200
  timeout 200 python tools/config_bench.py neg_shared skip-gram 8192 128 2 100
  timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 74 1000 ) 2>&1 | grep -v Warning | tee gpurun_out/s5a_configs.txt
echo "== gen-3 tests"; NNCF_EVAL_GEN=3 timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -15
echo "== g3 sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 g3 2>&1 | tee gpurun_out/s5a_sweep_g3.txt | tail -24
