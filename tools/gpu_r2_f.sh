#!/bin/bash
# round 2, call F (2 GPUs): multi-GPU parity under torchrun, N = 2 bench at the driver's settings, start-up probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02f_scale_n2_short.json 2> gpurun_out/r02f_scale_n2_short.err; echo "rc=$?"; tail -5 gpurun_out/r02f_scale_n2_short.err
timeout 600 $TR bench.py --gpus 2 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02f_scale_n2_long.json 2> gpurun_out/r02f_scale_n2_long.err; echo "rc=$?"; tail -3 gpurun_out/r02f_scale_n2_long.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-eval --cpu-steps 2 > gpurun_out/r02f_n1_short.json 2>/dev/null
python - <<PY
import json
for f in ("r02f_n1_short", "r02f_scale_n2_short", "r02f_scale_n2_long"):
    try:
        j=json.load(open("gpurun_out/%s.json" % f)); print(f, "value=%.3e ms/step=%.4f e2e=%.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["clocks"], j["arm"]["parallelism"][-90:])
    except Exception as ex: print(f, "ERR", ex)
PY
timeout 300 python tools/startup_probe.py
