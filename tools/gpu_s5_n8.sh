#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 1500 --warmup 100 > gpurun_out/r01c_scale_n8.json 2> gpurun_out/r01c_scale_n8.err; echo "rc=$?"; tail -2 gpurun_out/r01c_scale_n8.err; python - <<PY
import json
j=json.load(open("gpurun_out/r01c_scale_n8.json")); print("N=8 value=%.3e e2e=%.3e eval %s" % (j["value"], j["e2e"]["value"], j["extra"]["whole_at_k"]["users_per_sec"]))
PY
