#!/bin/bash
# round 2, call Y (1 GPU): whole GPU suite; zero-copy loss in the host-fed loop; bench N = 1 short window (default command)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02y_pytest.txt; cat gpurun_out/r02y_pytest.txt
for m in 0 1; do
  NNCF_HOST_LOSS_COPY=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02y_bench_copy$m.json 2> gpurun_out/r02y_bench.err
  python - <<PY
import json
j=json.load(open("gpurun_out/r02y_bench_copy$m.json"))
print("loss copy=$m: value %.3e  %.2f us/step   e2e %.3e  per_call %.3e" % (j["value"], j["ms_per_step"]*1e3, j["e2e"]["value"], j["e2e"]["per_call"]["value"]))
PY
done
timeout 900 python bench.py > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err; echo "default bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/r02y_bench_default.json"))
print("default: value %.3e  %.2f us/step  steps %d  e2e %.3e  clocks %s" % (j["value"], j["ms_per_step"]*1e3, j["steps"], j["e2e"]["value"], j["clocks"]))
print({k:(round(v["tflops"]),round(v["frac_of_bf16_burst"],3)) for k,v in j["extra"]["whole_at_k"]["by_k"].items()})
PY
