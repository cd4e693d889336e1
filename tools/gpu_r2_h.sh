#!/bin/bash
# round 2, call H: convergence table (40 epochs), ncu of the secondary kernels (summary only comes back)
mkdir -p gpurun_out
timeout 1500 python tools/convergence.py --epochs 40 --out gpurun_out/r02_convergence.md > gpurun_out/r02h_convergence.log 2>&1; tail -70 gpurun_out/r02h_convergence.log
KR='regex:sample_kernel|make_keys|radix_|scan_|group_emit|group_sample|pairs_|rows_sgd|adam_|meanpool_|score_pairs|eval_given'
timeout 1500 ncu --set full --clock-control none -k "$KR" -c 80 -o /tmp/r02_kernels python tools/kernel_zoo.py 1 > gpurun_out/r02_kernels.log 2>&1; echo "ncu rc=$?"
python tools/summarize_kernels.py /tmp/r02_kernels.ncu-rep gpurun_out/r02_kernels.log gpurun_out/r02_kernels_summary.md | head -60
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/r02_kernels_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
