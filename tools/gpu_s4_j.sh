#!/bin/bash
for cl in 1 2 4 8; do for dbg in 2 0; do echo "== CL=$cl DBG=$dbg"; NNCF_EVAL_CL=$cl NNCF_EVAL_DBG=$dbg timeout 120 python tools/eval_bench.py 36864 1000000 50; done; done
