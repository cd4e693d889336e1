#!/bin/bash
# session-4 evidence run: all GPU parity tests, the bench line, ncu launch lists of the training step and of whole@k
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python bench.py --steps 3000 --warmup 100 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r01b_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2000 --warmup 20 > gpurun_out/r01b_bench_reference.json 2>> gpurun_out/r01b_bench_n1.err; echo "ref rc=$?"
B="python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 3 --links 4000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather_rows|score_grad|finalize|loss_out' -s 150 -c 160 --csv \
   --log-file gpurun_out/r01b_launches_train.csv $B > gpurun_out/r01b_launches_train.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|rows_to_img|topk_merge' --csv \
   --log-file gpurun_out/r01b_launches_eval.csv python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r01b_launches_eval.log 2>&1
rm -f gpurun_out/r01b_timeline.txt
NNCF_TIMELINE=gpurun_out/r01b_timeline.txt timeout 600 python bench.py --steps 3000 --warmup 100 --no-eval --cpu-steps 20 > /dev/null 2>&1
python tools/timeline.py gpurun_out/r01b_timeline.txt 200 > gpurun_out/r01b_timeline_summary.txt 2>&1; cat gpurun_out/r01b_timeline_summary.txt
python - <<PY
import json
j=json.load(open("gpurun_out/r01b_bench_n1.json")); print("value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq=%.3e"%j["sequential"]["value"], "eval", j["extra"]["whole_at_k"]["users_per_sec"], j["extra"]["whole_at_k"]["roofline"]["frac"], "cpu", j["cpu_baseline"]["value"])
print(open("gpurun_out/r01b_bench_reference.json").read()[:300])
PY
