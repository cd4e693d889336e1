#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_eval_sampler_batch.py tests/test_gpu_end_to_end.py -m gpu -q -x 2>&1 | tail -8
for k in 50 10 100; do timeout 120 python tools/eval_bench.py 37888 1000000 $k; done
for dbg in 2 1 3; do echo "== NNCF_EVAL_DBG=$dbg"; NNCF_EVAL_DBG=$dbg timeout 120 python tools/eval_bench.py 37888 1000000 50; done
echo "== old kernel"; NNCF_EVAL_V1=1 timeout 120 python tools/eval_bench.py 37888 1000000 50
timeout 120 python tools/eval_bench.py 75776 2000000 50
