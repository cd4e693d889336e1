#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -4
CB="python tools/config_bench.py neg_shared skip-gram 512 128 37 2000 ureg"
echo "== 1M x 1M zipf 10,10 dedup"; ZIPF=10,10 timeout 120 $CB 2>&1 | tail -1
echo "== 1M x 1M zipf 10,10 nodedup"; ZIPF=10,10 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 hottest dedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 hottest nodedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== N=8 coolest dedup"; NU=125000 NI=62500 ZIPF=2.1,1.56 timeout 120 $CB 2>&1 | tail -1
echo "== N=2 dedup"; NU=500000 NI=250000 ZIPF=5,2.5 timeout 120 $CB 2>&1 | tail -1
echo "== R=1 dedup"; python tools/config_bench.py neg_shared skip-gram 512 128 1 3000 ureg 2>&1 | tail -1
