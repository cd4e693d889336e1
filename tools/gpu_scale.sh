#!/bin/bash
# usage: gpu_scale.sh N [extra bench args]  — bench on N GPUs (torchrun for N > 1)
N=$1; shift
mkdir -p gpurun_out
TAG=n$N$(echo "$*" | tr -d ' -')
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 2000 --warmup 100 "$@" > gpurun_out/scale_$TAG.json 2> gpurun_out/scale_$TAG.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
     bench.py --gpus $N --steps 2000 --warmup 100 "$@" > gpurun_out/scale_$TAG.json 2> gpurun_out/scale_$TAG.err
fi
echo "rc=$?"; tail -3 gpurun_out/scale_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/scale_$TAG.json"):
    l=l.strip()
    if l.startswith("{"):
        j=json.loads(l); print("N=%d value=%.3e ms/step=%.4f e2e=%.3e phases=%s" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"], {k: round(v,4) for k,v in j["roofline"]["phases_ms"].items()})); print(j["config"]["parallelism"])
PY
