#!/bin/bash
# ncu launch list + per-launch SOL/memory metrics of ONE timed step of bench.py, plus --set full source-level captures of
# the top kernels.  Big reports are reduced to CSV on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== bench (not under a profiler)"
timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "== ncu SOL+memory sections (one timed step, every launch)"
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
   --section Occupancy --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum \
   --clock-control none -o /tmp/step_full -f python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
ncu -i /tmp/step_full.ncu-rep --page raw --csv > /tmp/step_full_raw.csv 2>/dev/null
python scripts/ncu_reduce.py /tmp/step_full_raw.csv gpurun_out/step_full_metrics.csv
echo "== ncu --set full with source: conv kernels of one encoder"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'conv_gemm' -c 8 -o gpurun_out/top_conv -f \
   python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
echo "== ncu --set full with source: InfoNCE, EMA, BN kernels"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'infonce_main|ema_enqueue|bn_relu_maxpool|stem_pack|infonce_finalize' -c 7 -o gpurun_out/top_hbm -f \
   python bench.py --steps 1 --warmup 3 --profile-only >> gpurun_out/ncu_top.log 2>&1
tail -2 gpurun_out/ncu_top.log
for r in top_conv top_hbm; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > /tmp/$r.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/$r.csv gpurun_out/${r}_metrics.csv; done
du -sh gpurun_out; ls -la gpurun_out
