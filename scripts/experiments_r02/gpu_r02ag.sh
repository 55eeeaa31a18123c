#!/bin/bash
echo "== pytest backward"; timeout 1200 python -m pytest tests -m gpu -q -s -k "backward or train_step or fused_sgd or solver or loss_backward" 2>&1 | grep -E "gradients|losses|passed|failed|rror|assert|FAILED|solver|SGD" | tail -8
for c in 2 1; do
echo "== train step cfg$c"; timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['train_step']['ms_per_step'], d['train_step']['value'])"
done
