#!/bin/bash
# round 2, call U (1 GPU): tower graph-vs-eager diagnostic; bench N = 1 window with the long head-start sleep vs the old one; chunked host-fed loop
mkdir -p gpurun_out
timeout 600 python tools/tower_graph_diag.py > gpurun_out/r02u_tower_diag.txt 2>&1; tail -30 gpurun_out/r02u_tower_diag.txt
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q 2>&1 | tail -3
for sl in 200000 3000000; do for ch in 1 8; do
  NNCF_BENCH_SLEEP=$sl NNCF_HOST_CHUNK=$ch timeout 600 python bench.py --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02u_bench_sl${sl}_ch${ch}.json 2> gpurun_out/r02u_bench.err
  python - <<PY
import json
j=json.load(open("gpurun_out/r02u_bench_sl${sl}_ch${ch}.json"))
print("sleep $sl chunk $ch: value %.3e  %.2f us/step   e2e %.3e  per_call %.3e" % (j["value"], j["ms_per_step"]*1e3, j["e2e"]["value"], j["e2e"]["per_call"]["value"]))
PY
done; done
