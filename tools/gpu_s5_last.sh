#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3000 --warmup 100 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r01c_bench_n1.err
python - <<PY
import json
j=json.load(open("gpurun_out/r01c_bench_n1.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq %.3e"%j["sequential"]["value"], "adam %.3e"%j["sequential"]["lazy_adam"]["value"], "eval %.1f TF"%j["extra"]["whole_at_k"]["roofline"]["achieved"], "c5 %.3e"%j["sequential"]["c5_max_margin_b16384_d256"]["value"])
PY
