#!/bin/bash
# round 2, call O (8 GPUs): dedup in the fused drain + faithful block distributions: N = 8 / 4 / 1 short and long; tests first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_multi.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -4
run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n "$@"; fi; }
run 8 --steps 20 --warmup 5 --no-eval > gpurun_out/r02o_scale_n8.json 2> gpurun_out/r02o_scale_n8.err; echo "n8 rc=$?"
NNCF_DEDUP=0 run 8 --steps 20 --warmup 5 --no-eval > gpurun_out/r02o_scale_n8_nodedup.json 2> /dev/null; echo "n8 nodedup rc=$?"
run 4 --steps 20 --warmup 5 --no-eval > gpurun_out/r02o_scale_n4.json 2> gpurun_out/r02o_scale_n4.err; echo "n4 rc=$?"
run 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02o_scale_n2.json 2> gpurun_out/r02o_scale_n2.err; echo "n2 rc=$?"
run 1 --steps 20 --warmup 5 --no-eval --cpu-steps 2 > gpurun_out/r02o_scale_n1.json 2> gpurun_out/r02o_scale_n1.err; echo "n1 rc=$?"
run 8 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02o_scale_n8_long.json 2> gpurun_out/r02o_scale_n8_long.err; echo "n8 long rc=$?"
python - <<PY
import json
v1=None
for f in ("r02o_scale_n1","r02o_scale_n2","r02o_scale_n4","r02o_scale_n8","r02o_scale_n8_nodedup","r02o_scale_n8_long"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f))
        if f=="r02o_scale_n1": v1=j["value"]
        print(f, "N=%d value=%.3e ms/step=%.4f e2e=%.3e" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"]), "eff=%.3f" % (j["value"]/(j["n_gpus"]*v1)), {k:round(v*1e3,1) for k,v in j["roofline"]["phases_ms"].items()})
    except Exception as ex: print(f, "ERR", ex)
PY
