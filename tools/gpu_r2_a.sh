#!/bin/bash
# round 2, call A: baseline (tests, driver-style bench) + in-chain ablations of the score kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_short.json 2> gpurun_out/r02a_bench_short.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/r02a_bench_short.json")); print("short: value=%.3e"%j["value"], "ms/step=%.4f"%j["ms_per_step"], "e2e=%.3e"%j["e2e"]["value"], "seq %.3e"%j["sequential"]["value"], j["clocks"])
PY
CB="python tools/config_bench.py neg_shared skip-gram 512 128"
echo "== base"; timeout 120 $CB 37 3000 2>&1 | tail -1
for n in 1 2 4 8 16 32 64 17 3 12; do
  echo "== ablate $n"; NNCF_LIB_PATH=build/ablate/libnncf_ab$n.so timeout 120 $CB 37 3000 2>&1 | tail -1
done
echo "== base R=1"; timeout 120 $CB 1 3000 2>&1 | tail -1
echo "== base R=74"; timeout 120 $CB 74 2000 2>&1 | tail -1
echo "== adam R=37"; timeout 120 $CB 37 1000 adam 2>&1 | tail -1
echo "== adam R=1"; timeout 120 $CB 1 2000 adam 2>&1 | tail -1
NNCF_TIMELINE=gpurun_out/r02a_timeline.txt timeout 120 $CB 37 3000 2>&1 | tail -1
python tools/timeline.py gpurun_out/r02a_timeline.txt > gpurun_out/r02a_timeline_summary.txt 2>&1; cat gpurun_out/r02a_timeline_summary.txt
