#!/bin/bash
# final refresh after the InfoNCE tail fixes: kernel-only timing, ncu capture, host probe, bench lines
mkdir -p gpurun_out
python tests/nce_kernel_probe.py 2>&1 | tee gpurun_out/r02_nce_kernel_probe.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'infonce_main|ema_enqueue' -c 4 -o gpurun_out/r02_top_nce -f \
   python bench.py --config 2 --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
ncu -i gpurun_out/r02_top_nce.ncu-rep --page raw --csv > /tmp/top_nce.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/top_nce.csv gpurun_out/r02_top_nce_set_full.csv
rm -f gpurun_out/r02_top_nce.ncu-rep
python tests/nce_host_probe.py ResNet50 2>&1 | head -1 | tee gpurun_out/r02_nce_host_probe.log
python tests/nce_host_probe.py ResNet18 2>&1 | head -1 | tee -a gpurun_out/r02_nce_host_probe.log
for c in 2 1 4; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_cfg$c.err > gpurun_out/r02_bench_line_cfg$c.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg$c.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], 'train', t.get('ms_per_step'), 'cpu', d['cpu_baseline']['value'], 'refgpu', (d.get('reference_same_gpu') or {}).get('value'), d['clocks'])" || tail -5 gpurun_out/bench_cfg$c.err
done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -3
