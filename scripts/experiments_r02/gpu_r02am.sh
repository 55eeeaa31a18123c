#!/bin/bash
mkdir -p gpurun_out
for c in 1 2; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 --ref-gpu 2> gpurun_out/bench_cfg$c.err > gpurun_out/r02_bench_line_cfg$c.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg$c.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], d['warmup'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], 'train', t.get('ms_per_step'), 'cpu', d['cpu_baseline']['value'], 'refgpu', (d.get('reference_same_gpu') or {}).get('value'), d['clocks'])" || tail -5 gpurun_out/bench_cfg$c.err
done
