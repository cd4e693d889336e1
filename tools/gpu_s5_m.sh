#!/bin/bash
mkdir -p gpurun_out
echo "== tests (eval g2)"; timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -2
echo "== tests (eval g3)"; NNCF_EVAL_GEN=3 timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -2
echo "== sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 quick2 2>&1 | grep -v Warn | tee gpurun_out/s5m_sweep.txt | tail -40
NNCF_EVAL_GEN=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc3 -s 1 -c 1 -o gpurun_out/s5m_eval3 python tools/eval_bench.py 37888 1000000 50 > gpurun_out/s5m_eval3.log 2>&1
tail -2 gpurun_out/s5m_eval3.log
