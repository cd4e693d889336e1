#!/bin/bash
# session 5: full GPU parity suite + bench lines + launch lists for the round-1c state
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 3000 --warmup 100 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r01c_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2000 --warmup 20 > gpurun_out/r01c_bench_reference.json 2>> gpurun_out/r01c_bench_n1.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|rows_to_img|topk_merge' --csv \
   --log-file gpurun_out/r01c_launches_eval.csv python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r01c_launches_eval.log 2>&1
python - <<PY
import json
j=json.load(open("gpurun_out/r01c_bench_n1.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "eval", j["extra"]["whole_at_k"], "cpu", j["cpu_baseline"]["value"])
print(open("gpurun_out/r01c_bench_reference.json").read()[:300])
PY
