#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): full GPU test suite, smoke, bench lines for configs 2 / 1 / 4,
# ncu launch lists + per-launch DRAM / tensor-pipe metrics of ONE timed step of configs 2 and 1, a --set full capture of the
# convolution kernel, the launch list of one training step, and compute-sanitizer memcheck / racecheck of the smoke step.
# Big reports are reduced to CSV on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
for c in 2 1 4; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 --ref-gpu 2> gpurun_out/bench_cfg$c.err > gpurun_out/r02_bench_line_cfg$c.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg$c.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], 'train', t.get('ms_per_step'), 'cpu', d['cpu_baseline']['value'], d['clocks'])"
done
for c in 2 1; do
echo "== ncu launch list cfg$c"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
echo "== ncu per-launch metrics cfg$c"
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -o /tmp/step_full_$c -f python bench.py --config $c --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
ncu -i /tmp/step_full_$c.ncu-rep --page raw --csv > /tmp/step_full_raw_$c.csv 2>/dev/null
python scripts/ncu_reduce.py /tmp/step_full_raw_$c.csv gpurun_out/r02_step_metrics_cfg$c.csv
done
echo "== ncu launch list: training step cfg2"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_train_cfg2.csv python bench.py --config 2 --profile-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
echo "== ncu --set full with source: conv kernels (config 2)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'conv_gemm' -s 20 -c 8 -o gpurun_out/r02_top_conv -f \
   python bench.py --config 2 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
tail -1 gpurun_out/ncu_top.log
ncu -i gpurun_out/r02_top_conv.ncu-rep --page raw --csv > /tmp/top_conv.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/top_conv.csv gpurun_out/r02_top_conv_set_full.csv
echo "== ncu --set full: InfoNCE forward (single launch) + EMA/enqueue"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'infonce_main|ema_enqueue' -c 4 -o gpurun_out/r02_top_nce -f \
   python bench.py --config 2 --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
tail -1 gpurun_out/ncu_top.log
ncu -i gpurun_out/r02_top_nce.ncu-rep --page raw --csv > /tmp/top_nce.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/top_nce.csv gpurun_out/r02_top_nce_set_full.csv
rm -f gpurun_out/r02_top_nce.ncu-rep
echo "== host cost of the InfoNCE step"; python tests/nce_host_probe.py ResNet50 2>&1 | head -2 | tee gpurun_out/r02_nce_host_probe.log
python tests/nce_host_probe.py ResNet18 2>&1 | head -2 | tee -a gpurun_out/r02_nce_host_probe.log
echo "== per-layer conv timings"; { python tests/conv_bench.py --iters 7; echo "-- transposed statistics pass"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --tstats; echo "-- apply epilogue + residual planes"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1; } 2>&1 | tee gpurun_out/r02_conv_layers.log | tail -3
echo "== compute-sanitizer memcheck: smoke + small training step"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tests/train_probe.py ResNet18 8 64 2>&1 | grep -E "ERROR SUMMARY|GLOBAL|Invalid|Error" | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_train.log
echo "== compute-sanitizer racecheck: smoke"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_racecheck_smoke.log
du -sh gpurun_out; ls gpurun_out | grep r02
