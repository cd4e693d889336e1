#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 2000 --warmup 100 > gpurun_out/bench_r37.json 2> gpurun_out/bench_r37.err; echo "rc=$?"; tail -3 gpurun_out/bench_r37.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_r37.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "tc_frac=%.3f"%j["roofline"]["frac"], "eval", j["extra"]["whole_at_k"]["users_per_sec"], j["extra"]["whole_at_k"]["roofline"]["frac"], "cpu", j["cpu_baseline"]["value"])
PY
