#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py tests/test_gpu_sampling_trainers.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 1000 --warmup 100 --no-eval --cpu-steps 50 > gpurun_out/s5z_bench.json 2> gpurun_out/s5z_bench.err; tail -2 gpurun_out/s5z_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/s5z_bench.json")); print("value=%.3e"%j["value"], "seq", j["sequential"])
PY
