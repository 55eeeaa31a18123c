#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (training)"; timeout 1500 python -m pytest tests -m gpu -q -s -k "backward or train_step or fused_sgd" 2>&1 | grep -E "gradients|losses|passed|failed|rror|assert|FAILED|cfg0" | tail -10
for c in 1 2; do
echo "== bench cfg$c"; timeout 900 python bench.py --config $c --steps 10 --warmup 3 2> gpurun_out/bench_k$c.err > gpurun_out/bench_k$c.json; python -c "
import json
d=json.load(open('gpurun_out/bench_k$c.json')); t=d['train_step']; print(d['value'], d['ms_per_step'], d['clocks'], 'train', t.get('ms_per_step'), t.get('value'), t.get('error'))"
done
echo "== ncu launch list: one training step cfg2"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_train_cfg2.csv python bench.py --config 2 --profile-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
