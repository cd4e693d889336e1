#!/bin/bash
mkdir -p gpurun_out
for a in "512 1" "512 37"; do echo "== $a"; timeout 60 ./build/score_bench $a | grep -v "stamp 1[0-5]\|stamp 2[0-35-9]\|stamp 3"; done
NNCF_DUMP_CTAS=1 timeout 60 ./build/score_bench 512 37 | grep "cta " > gpurun_out/s4b_ctas.txt; head -40 gpurun_out/s4b_ctas.txt
timeout 900 python -m pytest tests/test_gpu_train_step.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 2000 --warmup 100 --no-eval --cpu-steps 20 > gpurun_out/s4b_bench.json 2> gpurun_out/s4b_bench.err; echo "rc=$?"; tail -2 gpurun_out/s4b_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s4b_bench.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "loss", j["final_loss"])
PY
