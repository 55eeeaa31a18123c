#!/bin/bash
mkdir -p gpurun_out
{
echo "== conv apply 1"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
echo "== conv apply 2"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 2
} 2>&1 | tee gpurun_out/conv_variants_ab2.log | grep -E "^==|layer1 1x1 64|layer2 1x1 128|layer3 1x1 256->1024"
for v in "1 1" "0 1" "1 1" "0 1" "1 2"; do set -- $v
echo "== bench --config 2 RES_FETCH=$1 TWOPASS=$2"; VINCE_B200_RES_FETCH=$1 VINCE_B200_TWOPASS=$2 timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
