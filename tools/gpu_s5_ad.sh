#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/s5ad_n2.json 2> gpurun_out/s5ad_n2.err; echo "rc=$?"
wc -l gpurun_out/s5ad_n2.json; head -c 120 gpurun_out/s5ad_n2.json; echo; grep -c "NCCL version" gpurun_out/s5ad_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 50 --warmup 5 > gpurun_out/s5ad_ref_n2.json 2> gpurun_out/s5ad_ref_n2.err; echo "rc=$?"; wc -l gpurun_out/s5ad_ref_n2.json; head -c 120 gpurun_out/s5ad_ref_n2.json; echo
