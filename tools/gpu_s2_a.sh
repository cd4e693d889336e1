#!/bin/bash
# session-2 baseline: GPU parity suite + default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 2000 --warmup 100 > gpurun_out/s2a_bench.json 2> gpurun_out/s2a_bench.err; echo "rc=$?"
tail -3 gpurun_out/s2a_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s2a_bench.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], j["extra"]["whole_at_k"]["users_per_sec"], j["extra"]["whole_at_k"]["roofline"]["frac"], j["cpu_baseline"])
PY
