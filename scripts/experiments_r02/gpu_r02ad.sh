#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|gradients|losses|cfg|FAILED|skipped|halves|solver|SGD steps|saturation|stats_only|transposed" | tee gpurun_out/r02_pytest_gpu.log | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
echo "== default bench (what the driver runs)"; timeout 900 python bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['steps'], d['warmup'], 'e2e', d['e2e']['value'], d['gpu_launches'])"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
