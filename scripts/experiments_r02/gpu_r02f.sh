#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (training)"; timeout 1500 python -m pytest tests -m gpu -q -s -k "backward or train_step or fused_sgd or loss_backward" 2>&1 | grep -E "gradients|losses|passed|failed|rror|assert|FAILED|cfg0" | tail -30 | tee gpurun_out/pytest_gpu_f.log
echo "== bench cfg1"; timeout 900 python bench.py --config 1 --steps 10 --warmup 3 2> gpurun_out/bench_f1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'train', d['train_step'])"
tail -3 gpurun_out/bench_f1.err
echo "== bench cfg2"; timeout 900 python bench.py --config 2 --steps 10 --warmup 3 2> gpurun_out/bench_f2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'train', d['train_step'])"
tail -3 gpurun_out/bench_f2.err
