#!/bin/bash
# round 2, call S (1 GPU): full GPU suite on the working tree, the driver's N = 1 bench line, dedup on / off on crowded ids
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02s_pytest.txt; cat gpurun_out/r02s_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"
CB="python tools/config_bench.py neg_shared skip-gram 512 128 37 2000 ureg"
{
echo "== 1M x 1M zipf 10,10 dedup"; ZIPF=10,10 timeout 120 $CB 2>&1 | tail -1
echo "== 1M x 1M zipf 10,10 nodedup"; ZIPF=10,10 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 hottest dedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 hottest nodedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 coolest dedup"; NU=125000 NI=62500 ZIPF=2.1,1.56 timeout 120 $CB 2>&1 | tail -1
echo "== N=2 dedup"; NU=500000 NI=250000 ZIPF=5,2.5 timeout 120 $CB 2>&1 | tail -1
echo "== R=1 dedup"; python tools/config_bench.py neg_shared skip-gram 512 128 1 3000 ureg 2>&1 | tail -1
} > gpurun_out/r02s_dedup.txt 2>&1
cat gpurun_out/r02s_dedup.txt
