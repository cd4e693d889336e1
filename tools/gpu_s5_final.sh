#!/bin/bash
# session 5 evidence run: all GPU parity tests, both bench arms, ncu launch lists of the training step and of whole@k
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 3000 --warmup 100 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r01c_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2000 --warmup 20 > gpurun_out/r01c_bench_reference.json 2>> gpurun_out/r01c_bench_n1.err; echo "ref rc=$?"
B="python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 3 --links 4000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather_rows|score_grad|finalize|loss_out' -s 150 -c 160 --csv \
   --log-file gpurun_out/r01c_launches_train.csv $B > gpurun_out/r01c_launches_train.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|rows_to_img|topk_merge' --csv \
   --log-file gpurun_out/r01c_launches_eval.csv python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r01c_launches_eval.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_grad_tc -s 60 -c 1 -o gpurun_out/r01c_score $B > gpurun_out/r01c_score.log 2>&1
rm -f gpurun_out/r01c_timeline.txt
NNCF_TIMELINE=gpurun_out/r01c_timeline.txt timeout 600 python bench.py --steps 3000 --warmup 100 --no-eval --cpu-steps 20 > /dev/null 2>&1
python tools/timeline.py gpurun_out/r01c_timeline.txt 200 > gpurun_out/r01c_timeline_summary.txt 2>&1; cat gpurun_out/r01c_timeline_summary.txt
python - <<PY
import json
j=json.load(open("gpurun_out/r01c_bench_n1.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq", j["sequential"], "eval", j["extra"]["whole_at_k"], "cpu", j["cpu_baseline"]["value"])
print(open("gpurun_out/r01c_bench_reference.json").read()[:300])
PY
