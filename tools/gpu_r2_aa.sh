#!/bin/bash
# round 2, call AA (1 GPU): symmetric G' exchange: parity, step times on / off; host-fed loop with the chunk ramp; short bench window
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k "g_exchange or host_fed" 2>&1 | tail -8
for gx in 0 1; do
  echo "=== NNCF_GX=$gx"
  for cfg in "neg_shared skip-gram 512 128 37 2000 ureg" "neg_shared mse 512 128 37 1000 ureg" "neg_shared skip-gram 512 64 37 1000 ureg" "neg_shared skip-gram 1024 128 9 1000 ureg" "neg_shared skip-gram 512 128 74 1000 ureg"; do
    NNCF_GX=$gx ZIPF=10,10 timeout 120 python tools/config_bench.py $cfg 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/r02aa_gx.txt
timeout 600 python tools/host_fed_bench.py 2>&1 | grep -v Warning | tee gpurun_out/r02aa_host_fed.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/r02aa_bench.json"))
print("bench 20/5: value %.3e  %.2f us/step   e2e %.3e  per_call %.3e" % (j["value"], j["ms_per_step"]*1e3, j["e2e"]["value"], j["e2e"]["per_call"]["value"]))
PY
