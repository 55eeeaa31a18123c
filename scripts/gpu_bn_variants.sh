#!/bin/bash
mkdir -p gpurun_out
for v in 41 42 44 81 82 84; do
  echo "== bn_apply variant $v"
  VINCE_B200_BNAPPLY=$v timeout 120 python tests/elem_bench.py --only bn_apply 2>&1 | tail -12
done | tee gpurun_out/bn_variants.log
bash scripts/gpu_check.sh
