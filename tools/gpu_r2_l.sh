#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_end_to_end.py -x -q -k meanpool_graph 2>&1 | tail -40
for sch in neg_shared group_neg_shared; do for gmode in 1 0; do timeout 300 python tools/tower_bench.py $sch $gmode 2>&1 | tail -4; done; done
CB="python tools/config_bench.py neg_shared"
echo "== C5"; timeout 120 $CB max-margin 16384 256 1 60 norm ureg 2>&1 | tail -1
echo "== C5 split1"; NNCF_SPLIT=1 timeout 120 $CB max-margin 16384 256 1 60 norm ureg 2>&1 | tail -1
echo "== B4096 R5"; timeout 120 $CB skip-gram 4096 128 5 300 ureg 2>&1 | tail -1
echo "== B4096 R5 split1"; NNCF_SPLIT=1 timeout 120 $CB skip-gram 4096 128 5 300 ureg 2>&1 | tail -1
echo "== R=18"; timeout 120 $CB skip-gram 512 128 18 3000 ureg 2>&1 | tail -1
echo "== R=9"; timeout 120 $CB skip-gram 512 128 9 3000 ureg 2>&1 | tail -1
