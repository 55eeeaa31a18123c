#!/bin/bash
# ncu launch list + --set full capture of ONE timed step of bench.py; the big .ncu-rep is reduced to CSV on the box
# (gpurun_out/ is capped at 64 MiB) and only a small source-level capture of the top kernels is kept as .ncu-rep.
mkdir -p gpurun_out
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "== ncu SOL+memory sections (one timed step, every launch)"
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
   --section Occupancy --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum \
   --clock-control none -o /tmp/step_full -f python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/step_full.ncu-rep --page raw --csv > /tmp/step_full_raw.csv 2>/dev/null
python scripts/ncu_reduce.py /tmp/step_full_raw.csv gpurun_out/step_full_metrics.csv
ls -la /tmp/step_full* gpurun_out/
echo "== ncu full with source: 4 conv launches + InfoNCE + EMA"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'conv_gemm|infonce_main|ema_enqueue' -s 4 -c 6 -o gpurun_out/top_kernels -f \
   python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
tail -2 gpurun_out/ncu_top.log
du -sh gpurun_out
