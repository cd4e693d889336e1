#!/bin/bash
# round 2, call E: new bench.py at the driver's settings (N = 1), C5 workload, reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_short.json 2> gpurun_out/r02e_bench_short.err ) 2>&1 | grep real; echo "rc=$?"; tail -3 gpurun_out/r02e_bench_short.err
python - <<PY
import json
j=json.load(open("gpurun_out/r02e_bench_short.json"))
print("short: value=%.3e ms/step=%.4f e2e=%.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["clocks"])
print(" roofline", {k: j["roofline"][k] for k in ("bound","achieved","frac","phases_ms")})
print(" ref_sem", j["config"].get("reference_semantics"))
print(" seq", {k:v for k,v in j["sequential"].items() if k in ("value","ms_per_step","batch_size_sweep")})
print(" c5", j["sequential"].get("c5_max_margin_b16384_d256"))
print(" eval", json.dumps(j["extra"]["whole_at_k"]["by_k"]))
print(" cpu", j["cpu_baseline"])
PY
( time timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-eval > gpurun_out/r02e_bench_c5.json 2> gpurun_out/r02e_bench_c5.err ) 2>&1 | grep real; tail -3 gpurun_out/r02e_bench_c5.err
python - <<PY
import json
j=json.load(open("gpurun_out/r02e_bench_c5.json"))
print("c5: value=%.3e ms/step=%.4f e2e=%.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["clocks"])
print(" roofline", {k: j["roofline"][k] for k in ("bound","achieved","frac","phases_ms")})
print(" cpu", j["cpu_baseline"])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02e_bench_reference.json 2>/dev/null; cat gpurun_out/r02e_bench_reference.json | cut -c1-300
