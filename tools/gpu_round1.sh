#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 2000 --warmup 100 > gpurun_out/bench_r32.json 2> gpurun_out/bench_r32.err; echo "rc=$?" >> gpurun_out/bench_r32.err
for R in 1 4 16 64 128; do
  timeout 300 python bench.py --steps 1000 --warmup 50 --replicas $R --no-eval --cpu-steps 20 --links 20000000 > gpurun_out/bench_r$R.json 2> gpurun_out/bench_r$R.err
done
tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_r32.json; tail -5 gpurun_out/bench_r32.err
for R in 1 4 16 64 128; do python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_r$R.json")); print("R=$R", "value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", j["roofline"]["phases_ms"], "e2e=%.3e"%j["e2e"]["value"])
except Exception as e: print("R=$R failed", e)
PY
done
