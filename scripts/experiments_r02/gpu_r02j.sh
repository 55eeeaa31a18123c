#!/bin/bash
mkdir -p gpurun_out
export PROBE_FP32_REF=1
python tests/train_probe.py ResNet18 32 96 2>&1 | grep -E "fp32 oracle|GLOBAL|loss"
VINCE_B200_WGRAD_PASSES=3 python tests/train_probe.py ResNet18 32 96 2>&1 | grep -E "GLOBAL"
python tests/train_probe.py ResNet50 16 96 2>&1 | grep -E "fp32 oracle|GLOBAL|loss"
echo "== pytest (training)"; timeout 1500 python -m pytest tests -m gpu -q -s -k "backward or train_step or fused_sgd" 2>&1 | grep -E "gradients|losses|passed|failed|rror|assert|FAILED|cfg0" | tail -10
for c in 1 2; do
echo "== bench cfg$c"; timeout 900 python bench.py --config $c --steps 10 --warmup 3 2> gpurun_out/bench_j$c.err > gpurun_out/bench_j$c.json; python -c "
import json
d=json.load(open('gpurun_out/bench_j$c.json')); print(d['value'], d['ms_per_step'], d['clocks'], 'train', d['train_step'])"
done
echo "== ncu launch list: one training step cfg1"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_train_cfg1.csv python bench.py --config 1 --profile-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
