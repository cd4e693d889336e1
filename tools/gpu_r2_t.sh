#!/bin/bash
# round 2, call T (1 GPU): the failing tower-graph test with its message, the whole GPU suite without -x, ncu --set full of the score kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_end_to_end.py -x -q -k meanpool_graph 2>&1 | grep -E "^E|Error|assert|passed|failed" | head -40 > gpurun_out/r02t_tower.txt; cat gpurun_out/r02t_tower.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02t_pytest.txt; cat gpurun_out/r02t_pytest.txt
CB="python tools/config_bench.py neg_shared skip-gram 512 128 37 60 ureg"
ZIPF=10,10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'score_grad_tc' -s 40 -c 1 -o gpurun_out/r02t_score $CB > gpurun_out/r02t_score.log 2>&1
ncu -i gpurun_out/r02t_score.ncu-rep --page source --csv > gpurun_out/r02t_score_source.csv 2>/dev/null
ncu -i gpurun_out/r02t_score.ncu-rep --page raw --csv > gpurun_out/r02t_score_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02t
