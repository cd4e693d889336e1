#!/bin/bash
# round 2, call N (8 GPUs): the driver's scaling run (N = 1, 2, 4, 8 at --steps 20 --warmup 5), C5 at 8 GPUs
mkdir -p gpurun_out
run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n "$@"; fi; }
run 8 --steps 20 --warmup 5 > gpurun_out/r02n_scale_n8.json 2> gpurun_out/r02n_scale_n8.err; echo "n8 rc=$?"; tail -2 gpurun_out/r02n_scale_n8.err
run 4 --steps 20 --warmup 5 --no-eval > gpurun_out/r02n_scale_n4.json 2> gpurun_out/r02n_scale_n4.err; echo "n4 rc=$?"
run 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02n_scale_n2.json 2> gpurun_out/r02n_scale_n2.err; echo "n2 rc=$?"
run 1 --steps 20 --warmup 5 --no-eval --cpu-steps 2 > gpurun_out/r02n_scale_n1.json 2> gpurun_out/r02n_scale_n1.err; echo "n1 rc=$?"
run 8 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02n_scale_n8_long.json 2> gpurun_out/r02n_scale_n8_long.err; echo "n8 long rc=$?"
run 8 --workload c5 --steps 20 --warmup 5 --no-eval > gpurun_out/r02n_c5_n8.json 2> gpurun_out/r02n_c5_n8.err; echo "c5 n8 rc=$?"; tail -2 gpurun_out/r02n_c5_n8.err
run 1 --workload c5 --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02n_c5_n1.json 2> gpurun_out/r02n_c5_n1.err; echo "c5 n1 rc=$?"
python - <<PY
import json
v1=None
for f in ("r02n_scale_n1","r02n_scale_n2","r02n_scale_n4","r02n_scale_n8","r02n_scale_n8_long","r02n_c5_n1","r02n_c5_n8"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f))
        if f=="r02n_scale_n1": v1=j["value"]
        print(f, "N=%d value=%.3e ms/step=%.4f e2e=%.3e" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"]), "eff=%.3f" % (j["value"]/(j["n_gpus"]*v1)) if v1 and "scale" in f else "", j["clocks"].get("sm_mhz"), j["clocks"].get("reasons"))
        if j.get("extra",{}).get("whole_at_k"): print("   eval", {k:(round(v["users_per_sec"]),round(v["tflops"],1)) for k,v in j["extra"]["whole_at_k"]["by_k"].items()})
    except Exception as ex: print(f, "ERR", ex)
PY
