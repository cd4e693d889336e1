#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/update_bench 37888 1.0
timeout 120 ./build/update_bench 37888 0.8
rm -f gpurun_out/s4c_timeline.txt
NNCF_TIMELINE=gpurun_out/s4c_timeline.txt timeout 600 python bench.py --steps 1500 --warmup 100 --no-eval --cpu-steps 20 > gpurun_out/s4c_bench.json 2> gpurun_out/s4c_bench.err; echo "rc=$?"; tail -2 gpurun_out/s4c_bench.err
python tools/timeline.py gpurun_out/s4c_timeline.txt 200 2>&1 | head -30
for R in 37 74 111 148; do
timeout 600 python bench.py --steps 1500 --warmup 100 --no-eval --cpu-steps 20 --replicas $R > gpurun_out/s4c_bench_R$R.json 2> gpurun_out/s4c_bench.err; echo "rc=$?"; tail -2 gpurun_out/s4c_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s4c_bench_R$R.json")); print("R=$R value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"])
PY
done
