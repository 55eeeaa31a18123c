#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_s.log
{
for st in 2 4; do
echo "== conv default STAGING=$st"; VINCE_B200_STAGING=$st python tests/conv_bench.py --filter "r50.layer" --iters 7
echo "== conv apply 1 STAGING=$st"; VINCE_B200_STAGING=$st python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
done
echo "== conv apply 0"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 0
} 2>&1 | tee gpurun_out/conv_variants_s.log
for e in 1 2 1 2; do
echo "== bench --config 1 EPI_SETS=$e"; VINCE_B200_EPI_SETS=$e timeout 600 python bench.py --config 1 --steps 30 --warmup 5 --profile-only 2>&1 | tail -1
done
for tp in 0 1 0 1; do
echo "== bench --config 2 TWOPASS=$tp"; VINCE_B200_TWOPASS=$tp timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
