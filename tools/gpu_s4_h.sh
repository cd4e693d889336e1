#!/bin/bash
for dbg in 2 1 3 0; do echo "== NNCF_EVAL_DBG=$dbg"; NNCF_EVAL_DBG=$dbg timeout 120 python tools/eval_bench.py 37888 1000000 50; done
for k in 10 100; do NNCF_EVAL_DBG=0 timeout 120 python tools/eval_bench.py 37888 1000000 $k; done
