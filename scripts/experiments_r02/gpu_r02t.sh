#!/bin/bash
mkdir -p gpurun_out
for tp in 1 2 1 2; do
echo "== bench --config 2 TWOPASS=$tp"; VINCE_B200_TWOPASS=$tp timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tests/nce_host_probe.py ResNet18 2>&1 | head -2
