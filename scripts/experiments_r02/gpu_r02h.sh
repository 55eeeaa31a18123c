#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"; python tests/conv_bench.py --filter "r50." --iters 7
echo "== SMALLK_PAIR=1"; VINCE_B200_SMALLK_PAIR=1 python tests/conv_bench.py --filter "r50." --iters 7
echo "== SMALLK_PAIR=1 bn128"; VINCE_B200_SMALLK_PAIR=1 python tests/conv_bench.py --filter "r50." --iters 7 --bn 128
echo "== bn128"; python tests/conv_bench.py --filter "r50." --iters 7 --bn 128
} 2>&1 | tee gpurun_out/conv_variants_h.log
for c in 1 2; do
echo "== bench cfg$c"; timeout 900 python bench.py --config $c --steps 10 --warmup 3 2> gpurun_out/bench_h$c.err > gpurun_out/bench_h$c.json; python -c "
import json
d=json.load(open('gpurun_out/bench_h$c.json')); print(d['value'], d['ms_per_step'], d['clocks'], 'train', d['train_step']['ms_per_step'] if d['train_step'] else None, d['train_step'].get('gpu_launches_per_step'))"
done
echo "== pytest (training)"; timeout 1500 python -m pytest tests -m gpu -q -s -k "train_step or fused_sgd" 2>&1 | grep -E "gradients|losses|passed|failed|rror|assert|FAILED|cfg0" | tail -10
