#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_cfg2.err > gpurun_out/r02_bench_line_cfg2.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_cfg2.json')); t=d['train_step'] or {}
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], 'nce', d['infonce_step_ms'], 'train', t.get('ms_per_step'), 'cpu', d['cpu_baseline']['value'], d['clocks'])"
