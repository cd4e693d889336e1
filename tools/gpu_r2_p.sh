#!/bin/bash
# round 2, call P: why does the score kernel slow down on a stratified block (small tables, crowded ids)?
CB="python tools/config_bench.py neg_shared skip-gram 512 128 37 2000 ureg"
echo "== 1M x 1M uniform"; timeout 120 $CB 2>&1 | tail -1
echo "== 1M x 1M zipf 10,10"; ZIPF=10,10 timeout 120 $CB 2>&1 | tail -1
echo "== 125k x 62.5k uniform"; NU=125000 NI=62500 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== 125k x 62.5k zipf 1.25,0.625 (N=8 block, hottest stratum) nodedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== same, dedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 NNCF_DEDUP=1 timeout 120 $CB 2>&1 | tail -1
echo "== 125k x 62.5k zipf 2.1,1.56 (N=8 block, coolest stratum) nodedup"; NU=125000 NI=62500 ZIPF=2.1,1.56 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== same, no update (ablate via NNCF_FUSE_SGD=0 -> finalize path)"; NU=125000 NI=62500 ZIPF=2.1,1.56 NNCF_DEDUP=0 NNCF_FUSE_SGD=0 timeout 120 $CB 2>&1 | tail -1
echo "== 500k x 250k zipf 5,2.5 (N=2 block) nodedup"; NU=500000 NI=250000 ZIPF=5,2.5 NNCF_DEDUP=0 timeout 120 $CB 2>&1 | tail -1
echo "== same dedup"; NU=500000 NI=250000 ZIPF=5,2.5 NNCF_DEDUP=1 timeout 120 $CB 2>&1 | tail -1
