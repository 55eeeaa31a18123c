#!/bin/bash
mkdir -p gpurun_out
echo "== cfg1"; timeout 600 python bench.py --config 1 --steps 20 --warmup 5 --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['warmup'], 'e2e', d['e2e']['value'], d['clocks'])"
echo "== default bench"; timeout 600 python bench.py 2> gpurun_out/bench_cfg2.err > gpurun_out/r02_bench_line_cfg2.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg2.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'nce', d['infonce_step_ms'], 'train', t.get('ms_per_step'), d['clocks'])" || tail -5 gpurun_out/bench_cfg2.err
