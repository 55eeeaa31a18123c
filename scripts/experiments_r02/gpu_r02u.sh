#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -12
for c in 2 1; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_cfg$c.err > gpurun_out/r02_bench_line_cfg$c.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg$c.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], 'train', t.get('ms_per_step'), t.get('error'), 'cpu', d['cpu_baseline']['value'], 'refgpu', (d.get('reference_same_gpu') or {}).get('value'), d['clocks'])" || tail -5 gpurun_out/bench_cfg$c.err
done
for g in 0 1; do
echo "== train step GRAPH=$g"; VINCE_B200_GRAPH=$g timeout 600 python bench.py --config 2 --steps 5 --warmup 3 --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['train_step'])"
done
