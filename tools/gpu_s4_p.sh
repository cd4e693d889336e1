#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/s4p_bench.json 2> gpurun_out/s4p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s4p_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s4p_bench.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "eval", j["extra"]["whole_at_k"], "clocks", j["clocks"], "launches", j["gpu_launches"])
PY
