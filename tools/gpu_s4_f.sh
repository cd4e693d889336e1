#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/s4f_timeline.txt
NNCF_TIMELINE=gpurun_out/s4f_timeline.txt timeout 600 python bench.py --steps 20000 --warmup 100 --no-eval --cpu-steps 20 > gpurun_out/s4f_bench.json 2> gpurun_out/s4f_bench.err; echo "rc=$?"; tail -2 gpurun_out/s4f_bench.err
python tools/timeline.py gpurun_out/s4f_timeline.txt 200 2>&1 | head -30
python - <<PY
import json
j=json.load(open("gpurun_out/s4f_bench.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], j["clocks"])
PY
