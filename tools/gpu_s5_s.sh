#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1500 --warmup 100 > gpurun_out/r01c_scale_n2.json 2> gpurun_out/r01c_scale_n2.err; echo "rc=$?"; tail -3 gpurun_out/r01c_scale_n2.err; cut -c1-600 gpurun_out/r01c_scale_n2.json
