#!/bin/bash
# Round-2 evidence run (1 GPU): GPU suite, bench lines (default + the driver's short window + reference arm), ncu launch list
# of the bench command, ncu --set full of the dominant kernels, in-kernel timeline.  Outputs: gpurun_out/r02p_*; summarised
# into profiles/ by tools/summarize_profiles.py r02p.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02p_pytest.txt; cat gpurun_out/r02p_pytest.txt
timeout 900 python bench.py > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; echo "bench default rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench_n1_short.json 2> gpurun_out/r02p_bench_n1_short.err; echo "bench short rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02p_bench_reference.json 2> gpurun_out/r02p_bench_reference.err; echo "bench reference rc=$?"
B="python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 1 --links 4000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather_rows|score_grad|finalize|loss_out' -s 150 -c 160 --csv \
   --log-file gpurun_out/r02p_launches_train.csv $B > gpurun_out/r02p_launches_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_grad_tc' -s 30 -c 1 -o gpurun_out/r02p_score $B > gpurun_out/r02p_score.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'gather_rows_vec' -s 30 -c 1 -o gpurun_out/r02p_gather $B > gpurun_out/r02p_gather.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/r02p_eval python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r02p_eval.log 2>&1
for n in score gather eval; do ncu -i gpurun_out/r02p_$n.ncu-rep --page raw --csv > gpurun_out/r02p_${n}_raw.csv 2>/dev/null; done
NNCF_TIMELINE=gpurun_out/r02p_timeline.txt ZIPF=10,10 timeout 120 python tools/config_bench.py neg_shared skip-gram 512 128 37 3000 ureg 2>&1 | tail -1
python tools/timeline.py gpurun_out/r02p_timeline.txt > gpurun_out/r02p_timeline_summary.txt 2>&1; cat gpurun_out/r02p_timeline_summary.txt
rm -f gpurun_out/r02p_timeline.txt gpurun_out/r02p_gather.ncu-rep
python - <<PY
import json
for f in ("r02p_bench_n1","r02p_bench_n1_short","r02p_bench_reference"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.3e  %.2f us/step  e2e %.3e" % (j["value"], j["ms_per_step"]*1e3, j["e2e"]["value"]), j.get("clocks"))
        ex=j.get("extra",{})
        if "whole_at_k" in ex: print("   eval", {k:(round(v["users_per_sec"]),round(v["tflops"],1)) for k,v in ex["whole_at_k"]["by_k"].items()})
        if "content_tower" in ex: print("   tower", json.dumps(ex["content_tower"])[:900])
    except Exception as ex: print(f, "ERR", ex)
PY
