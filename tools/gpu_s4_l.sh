#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_eval_sampler_batch.py -m gpu -q -x 2>&1 | tail -4
for pf in 0 4 8 16; do echo "== PF=$pf"; NNCF_EVAL_PF=$pf timeout 120 python tools/eval_bench.py 37888 1000000 50; NNCF_EVAL_PF=$pf NNCF_EVAL_DBG=2 timeout 120 python tools/eval_bench.py 37888 1000000 50; done
for k in 10 100; do timeout 120 python tools/eval_bench.py 37888 1000000 $k; done
timeout 120 python tools/eval_bench.py 75776 2000000 50
