#!/bin/bash
# Round-2 A/B experiments (one documented runner instead of a scratch script per gpurun call).
#   usage (under gpurun):  bash tools/gpu_experiments_r02.sh <name> [...]
#   dedup    duplicate folding in the fused drain on / off, on the id distributions of 1 / 2 / 8-GPU stratified blocks
#   teams    4 x 32-column S' buffers vs 2 x 64 (needs build/ablate/libnncf_teams.so: nvcc -DNNCF_SCORE_TEAMS=1 of score_tc_nsub{1,2}.cu)
#   gx       symmetric G' exchange (two-sided score kernel) on / off
#   hostfed  host-fed loop vs device-fed loop, per chunk length
#   tower    graph-captured content-tower step vs the eager step: where the two end up (diagnostic), and the step time
#   suite2   the whole GPU suite twice (flakiness check)
#   tower_ncu  ncu launch list of the graph-captured content-tower step (which nodes the 190 us are)
#   mufu     issue cost of the special-function and packed-math instructions (tools/mufu_bench.cu -> build/probe/mufu_bench)
#   scale2   bench.py at N = 1 and N = 2 (needs gpurun --gpus 2), the driver's short window and a long one
#   scale8   bench.py at N = 8 / 4 / 2 / 1 the way the driver launches it (needs gpurun --gpus 8), C5 at N = 8
#   scale4   the same at N = 4 / 2 / 1 (gpurun --gpus 4)
#   ncu_c5   ncu --set full of the C5 step's gather and finalize kernels (B = 16,384, d = 256, l2-normalised rows, max-margin)
#   ncu_adam   ncu --set full of the lazy-Adam apply and the finalize kernel at R = 37, launch list of the C2-like chain at R = 1
#   ncu_score   ncu --set full + source page of the score kernel
mkdir -p gpurun_out
CB="python tools/config_bench.py"
C3="neg_shared skip-gram 512 128 37 2000 ureg"
run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n "$@"; fi; }
for what in "$@"; do
  echo "######## $what"
  case $what in
    dedup)
      { echo "== 1M x 1M zipf 10,10 dedup"; ZIPF=10,10 timeout 120 $CB $C3 2>&1 | tail -1
        echo "== 1M x 1M zipf 10,10 nodedup"; ZIPF=10,10 NNCF_DEDUP=0 timeout 120 $CB $C3 2>&1 | tail -1
        echo "== N=8 hottest stratum dedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 timeout 120 $CB $C3 2>&1 | tail -1
        echo "== N=8 hottest stratum nodedup"; NU=125000 NI=62500 ZIPF=1.25,0.625 NNCF_DEDUP=0 timeout 120 $CB $C3 2>&1 | tail -1
        echo "== N=8 coolest stratum dedup"; NU=125000 NI=62500 ZIPF=2.1,1.56 timeout 120 $CB $C3 2>&1 | tail -1
        echo "== N=2 block dedup"; NU=500000 NI=250000 ZIPF=5,2.5 timeout 120 $CB $C3 2>&1 | tail -1
      } | tee gpurun_out/r02_dedup.txt ;;
    teams)
      for lib in "" build/ablate/libnncf_teams.so; do
        echo "=== lib ${lib:-default (2 x 64)}"
        for cfg in "$C3" "neg_shared skip-gram 512 128 1 3000 ureg" "neg_shared skip-gram 4096 128 5 500 ureg" "neg_shared skip-gram 512 64 37 1000 ureg"; do
          NNCF_LIB_PATH=$lib ZIPF=10,10 timeout 120 $CB $cfg 2>&1 | tail -1
        done
      done | tee gpurun_out/r02_teams.txt ;;
    gx)
      timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k "g_exchange" 2>&1 | tail -3
      for gx in 0 1; do
        echo "=== NNCF_GX=$gx"
        for cfg in "$C3" "neg_shared mse 512 128 37 1000 ureg" "neg_shared skip-gram 512 64 37 1000 ureg" "neg_shared skip-gram 512 128 74 1000 ureg"; do
          NNCF_GX=$gx ZIPF=10,10 timeout 120 $CB $cfg 2>&1 | tail -1
        done
      done | tee gpurun_out/r02_gx.txt ;;
    hostfed) timeout 600 python tools/host_fed_bench.py 2>&1 | grep -v Warning | tee gpurun_out/r02_host_fed.txt ;;
    tower)
      timeout 600 python tools/tower_graph_diag.py 2>&1 | tail -30 | tee gpurun_out/r02_tower_diag.txt
      for s in neg_shared group_neg_shared; do for g in 1 0; do timeout 300 python tools/tower_bench.py $s $g 2>&1 | tail -1; done; done | tee gpurun_out/r02_tower_bench.txt ;;
    suite2)
      for i in 1 2; do timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4; done | tee gpurun_out/r02_suite2.txt ;;
    tower_ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/r02_tower_launches.csv python tools/tower_bench.py neg_shared 1 > gpurun_out/r02_tower_ncu.log 2>&1
      python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_tower_launches.csv")) if len(r) > 5]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ik].split("(")[0][:90], []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("launches %d, total %.1f us" % (len(rows) - 1, tot / 1e3))
for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])): print("%-92s %4d  avg %7.2f us  share %.3f" % (n, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
PY
      ;;
    mufu) ./build/probe/mufu_bench | tee gpurun_out/r02_mufu.txt ;;
    scale2)
      run 1 --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
      run 2 --steps 20 --warmup 5 --no-eval > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
      run 2 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02_scale_n2_long.json 2> gpurun_out/r02_scale_n2_long.err
      python - <<PY
import json
v1 = None
for f in ("r02_scale_n1", "r02_scale_n2", "r02_scale_n2_long"):
    j = json.load(open("gpurun_out/%s.json" % f))
    v1 = v1 or j["value"]
    print(f, "N=%d value=%.3e us/step=%.2f e2e=%.3e eff=%.3f" % (j["n_gpus"], j["value"], j["ms_per_step"] * 1e3, j["e2e"]["value"], j["value"] / (j["n_gpus"] * v1)))
PY
      ;;
    scale8)
      run 8 --steps 20 --warmup 5 --cpu-steps 1 > gpurun_out/r02q_scale_n8.json 2> gpurun_out/r02q_scale_n8.err; echo "n8 rc=$?"
      run 1 --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02q_scale_n1.json 2> gpurun_out/r02q_scale_n1.err; echo "n1 rc=$?"
      run 8 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02q_scale_n8_long.json 2> gpurun_out/r02q_scale_n8_long.err; echo "n8 long rc=$?"
      run 8 --workload c5 --steps 20 --warmup 5 --no-eval > gpurun_out/r02q_c5_n8.json 2> gpurun_out/r02q_c5_n8.err; echo "c5 n8 rc=$?"
      python - <<PY
import json
v1 = None
for f in ("r02q_scale_n1", "r02q_scale_n2", "r02q_scale_n4", "r02q_scale_n8", "r02q_scale_n8_long", "r02q_c5_n1", "r02q_c5_n8"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
        if f == "r02q_scale_n1": v1 = j["value"]
        print(f, "N=%d value=%.3e us/step=%.2f e2e=%.3e" % (j["n_gpus"], j["value"], j["ms_per_step"] * 1e3, j["e2e"]["value"]),
              ("eff=%.3f" % (j["value"] / (j["n_gpus"] * v1))) if v1 and "scale" in f else "", {k: round(v * 1e3, 1) for k, v in j["roofline"]["phases_ms"].items()}, j["clocks"].get("reasons"))
        w = j.get("extra", {}).get("whole_at_k")
        if w: print("   whole@k", {k: (round(v["users_per_sec"]), round(v["tflops"], 1)) for k, v in w["by_k"].items()})
    except Exception as ex: print(f, "ERR", ex)
PY
      ;;
    scale4)
      for n in 4 2 1; do
        run $n --steps 20 --warmup 5 --no-eval --cpu-steps 1 > gpurun_out/r02q_scale_n$n.json 2> gpurun_out/r02q_scale_n$n.err; echo "n$n rc=$?"
      done
      run 4 --steps 2000 --warmup 50 --no-eval > gpurun_out/r02q_scale_n4_long.json 2> gpurun_out/r02q_scale_n4_long.err; echo "n4 long rc=$?"
      python - <<PY
import json
v1 = json.load(open("gpurun_out/r02q_scale_n1.json"))["value"]
for f in ("r02q_scale_n1", "r02q_scale_n2", "r02q_scale_n4", "r02q_scale_n4_long"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
        print(f, "N=%d value=%.3e us/step=%.2f e2e=%.3e eff=%.3f" % (j["n_gpus"], j["value"], j["ms_per_step"] * 1e3, j["e2e"]["value"], j["value"] / (j["n_gpus"] * v1)),
              {k: round(v * 1e3, 1) for k, v in j["roofline"]["phases_ms"].items()}, j["clocks"].get("reasons"))
    except Exception as ex: print(f, "ERR", ex)
PY
      ;;
    ncu_c5)
      for k in gather_rows_vec finalize_vec pos_score; do
        timeout 600 ncu --set full --clock-control none -k regex:$k -s 6 -c 1 -o gpurun_out/r02_c5_$k $CB neg_shared max-margin 16384 256 1 20 ureg norm > gpurun_out/r02_c5_$k.log 2>&1
        ncu -i gpurun_out/r02_c5_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for k in ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active']:
    if k in h: print('%-64s %s %s' % (k, v[h.index(k)][:80], u[h.index(k)]))
"
        rm -f gpurun_out/r02_c5_$k.ncu-rep
      done | tee gpurun_out/r02_c5_ncu.txt
      for sp in 0 1 2 4; do echo "== NNCF_SPLIT=$sp (0 = the host's choice)"; NNCF_SPLIT=$sp timeout 120 $CB neg_shared max-margin 16384 256 1 200 ureg norm 2>&1 | tail -1; done | tee -a gpurun_out/r02_c5_ncu.txt
      echo "== u_reg = 0"; timeout 120 $CB neg_shared max-margin 16384 256 1 200 norm 2>&1 | tail -1 ;;
    ncu_adam)
      for k in adam_apply_accum finalize_vec; do
        cfg="neg_shared skip-gram 512 128 37 30 ureg adam"; [ $k = finalize_vec ] && cfg="group_neg_shared log-loss 512 128 37 30 ureg norm adam"
        ZIPF=10,10 timeout 300 ncu --set full --clock-control none -k regex:$k -s 20 -c 1 -o gpurun_out/r02_$k $CB $cfg > /dev/null 2>&1
        ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for k in ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active']:
    if k in h: print('%-64s %s %s' % (k, v[h.index(k)][:90], u[h.index(k)]))
"
        rm -f gpurun_out/r02_$k.ncu-rep
      done | tee gpurun_out/r02_adam_ncu.txt
      ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 24 --csv --log-file gpurun_out/r02_c2_launches_after.csv $CB group_neg_shared log-loss 512 128 1 100 ureg norm adam > /dev/null 2>&1
      python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_c2_launches_after.csv")) if len(r) > 5]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[1:17]: print("%-70s %8.2f us" % (r[ik].split("(")[0][:70], float(r[iv].replace(",", "")) / 1e3))
PY
      ;;
    ncu_score)
      ZIPF=10,10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'score_grad_tc' -s 40 -c 1 -o gpurun_out/r02_score $CB neg_shared skip-gram 512 128 37 60 ureg > gpurun_out/r02_score.log 2>&1
      ncu -i gpurun_out/r02_score.ncu-rep --page source --csv > gpurun_out/r02_score_source.csv 2>/dev/null
      ncu -i gpurun_out/r02_score.ncu-rep --page raw --csv > gpurun_out/r02_score_raw.csv 2>/dev/null ;;
    *) echo "unknown experiment $what" ;;
  esac
done
