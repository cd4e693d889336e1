#!/bin/bash
# round 2, call X (2 GPUs): default kernel restored (G' exchange compile-time, opt-in): parity + step times; N = 1 / 2 short windows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q 2>&1 | tail -3
for gx in 0 1; do
  echo "=== NNCF_GX=$gx"
  for cfg in "neg_shared skip-gram 512 128 37 2000 ureg" "neg_shared skip-gram 4096 128 5 500 ureg"; do
    NNCF_GX=$gx ZIPF=10,10 timeout 120 python tools/config_bench.py $cfg 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/r02x_gx.txt
./build/probe/mufu_bench > gpurun_out/r02x_mufu.txt 2>&1; cat gpurun_out/r02x_mufu.txt
run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n "$@"; fi; }
run 1 --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02x_scale_n1.json 2> gpurun_out/r02x_scale_n1.err; echo "n1 rc=$?"
run 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02x_scale_n2.json 2> gpurun_out/r02x_scale_n2.err; echo "n2 rc=$?"
run 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02x_scale_n2b.json 2> gpurun_out/r02x_scale_n2b.err; echo "n2b rc=$?"
run 2 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02x_scale_n2_long.json 2> gpurun_out/r02x_scale_n2_long.err; echo "n2 long rc=$?"
python - <<PY
import json
v1=None
for f in ("r02x_scale_n1","r02x_scale_n2","r02x_scale_n2b","r02x_scale_n2_long"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f))
        if f=="r02x_scale_n1": v1=j["value"]
        print(f, "N=%d value=%.3e ms/step=%.4f e2e=%.3e" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"]), "eff=%.3f" % (j["value"]/(j["n_gpus"]*v1)), {k:round(v*1e3,1) for k,v in j["roofline"]["phases_ms"].items()})
    except Exception as ex: print(f, "ERR", ex)
PY
