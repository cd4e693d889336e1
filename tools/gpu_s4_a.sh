#!/bin/bash
# session 4 validation: GPU parity tests, the bench line, launch list of the same command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
timeout 900 python bench.py --steps 2000 --warmup 100 --cpu-steps 1000 > gpurun_out/s4a_bench.json 2> gpurun_out/s4a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s4a_bench.err
cat gpurun_out/s4a_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
   --log-file gpurun_out/s4a_launches_train.csv python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 3 --links 4000000 > gpurun_out/s4a_launches_train.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/s4a_launches_train.csv | cut -c1-300
