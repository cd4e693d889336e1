#!/bin/bash
mkdir -p gpurun_out
for f in 0 1; do
NNCF_FUSE_SGD=$f timeout 600 python bench.py --steps 2000 --warmup 100 --no-eval --cpu-steps 20 > gpurun_out/s2e_bench_f$f.json 2> gpurun_out/s2e_bench_f$f.err; echo "rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/s2e_bench_f$f.json")); print("fuse=$f value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "phases", {k:round(v,4) for k,v in j["roofline"]["phases_ms"].items()}, "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"])
PY
done
