#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -4
echo "== memcheck apply routes"; timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tests/apply_probe.py 2>&1 | grep -E "ERROR SUMMARY|train=|Invalid|Error" | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_apply.log
echo "== racecheck apply routes"; timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python tests/apply_probe.py 2>&1 | grep -E "RACECHECK SUMMARY|train=|Race reported|and Read|and Write" | tail -12 | tee gpurun_out/r02_sanitizer_racecheck_apply.log
echo "== per-layer"; { python tests/conv_bench.py --iters 7; echo "-- transposed statistics pass"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --tstats; echo "-- apply epilogue + residual planes"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1; echo "-- apply epilogue + bn(raw) residual"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 2; echo "-- apply epilogue, no residual"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 0; } > gpurun_out/r02_conv_layers.log 2>&1
for c in 2 1 4; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_cfg$c.err > gpurun_out/r02_bench_line_cfg$c.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg$c.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], 'train', t.get('ms_per_step'), 'cpu', d['cpu_baseline']['value'], 'refgpu', (d.get('reference_same_gpu') or {}).get('value'), d['clocks'])" || tail -5 gpurun_out/bench_cfg$c.err
done
for c in 2; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -o /tmp/step_full_$c -f python bench.py --config $c --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/step_full_$c.ncu-rep --page raw --csv > /tmp/step_full_raw_$c.csv 2>/dev/null
python scripts/ncu_reduce.py /tmp/step_full_raw_$c.csv gpurun_out/r02_step_metrics_cfg$c.csv
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'conv_gemm' -s 20 -c 8 -o /tmp/r02_top_conv -f \
   python bench.py --config 2 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
ncu -i /tmp/r02_top_conv.ncu-rep --page raw --csv > /tmp/top_conv.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/top_conv.csv gpurun_out/r02_top_conv_set_full.csv
