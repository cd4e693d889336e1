#!/bin/bash
mkdir -p gpurun_out
for a in "512 1" "512 37" "2048 4"; do timeout 60 ./build/score_bench $a | grep "per launch"; done
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
timeout 600 python bench.py --steps 2000 --warmup 100 --no-eval --cpu-steps 20 > gpurun_out/s2h_bench.json 2> gpurun_out/s2h_bench.err; echo "rc=$?"; tail -2 gpurun_out/s2h_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s2h_bench.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "per_call=%.3e"%j["e2e"]["per_call"]["value"], "seq=%.3e"%j["sequential"]["value"], "loss", j["final_loss"])
PY
