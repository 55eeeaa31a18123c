#!/bin/bash
mkdir -p gpurun_out
{
echo "== probes"; timeout 600 python tests/gpu_probe.py 2>&1 | grep -v PASS | tail -12
echo "== stem"; timeout 120 python tests/elem_bench.py --only stem
echo "== nce"; timeout 120 python tests/elem_bench.py --only nce
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nce ms',d['infonce_step_ms'],'iso ms',d['roofline']['isolated']['ms_per_step'],'iso frac',d['roofline']['isolated']['frac'],'clk',d['clocks'])"
tail -3 gpurun_out/bench.err
} 2>&1 | tee gpurun_out/exp.log
