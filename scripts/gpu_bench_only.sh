#!/bin/bash
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'iso ms',d['roofline']['isolated']['ms_per_step'],'iso frac',d['roofline']['isolated']['frac'],'clk',d['clocks'])"
tail -5 gpurun_out/bench.err
