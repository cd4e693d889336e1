#!/bin/bash
mkdir -p gpurun_out
for a in "512 1" "512 32" "8192 1"; do timeout 60 ./build/score_bench $a | head -4; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 2000 --warmup 100 > gpurun_out/bench_r37.json 2> gpurun_out/bench_r37.err; echo "rc=$?" >> gpurun_out/bench_r37.err
for R in 1 18 74; do
  timeout 300 python bench.py --steps 1000 --warmup 50 --replicas $R --no-eval --cpu-steps 20 --links 20000000 > gpurun_out/bench_r$R.json 2> gpurun_out/bench_r$R.err
done
tail -5 gpurun_out/bench_r37.err
for R in 1 18 37 74; do python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_r$R.json")); print("R=$R", "value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "tc_frac=%.3f"%j["roofline"]["frac"], j.get("extra",{}).get("whole_at_k",{}).get("users_per_sec"), j.get("extra",{}).get("whole_at_k",{}).get("roofline",{}).get("frac"))
except Exception as e: print("R=$R failed", e)
PY
done
