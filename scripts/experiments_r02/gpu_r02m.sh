#!/bin/bash
# TMA residual ring in the apply epilogue, fused InfoNCE tail, CUDA graphs: parity + timings
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "rel|passed|failed|rror|assert|FAILED|stats_only|transposed" | tail -40 | tee gpurun_out/pytest_gpu_m.log
{
for r in 2 3 4; do
echo "== conv apply 1 RES_TILES=$r";   VINCE_B200_RES_TILES=$r python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
done
echo "== conv apply 2";   python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 2
echo "== conv apply 0";   python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 0
} 2>&1 | tee gpurun_out/conv_variants_m.log
for v in "1 0" "1 1"; do set -- $v
echo "== bench --config 2 GRAPH=$1 TWOPASS=$2"; VINCE_B200_GRAPH=$1 VINCE_B200_TWOPASS=$2 timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --no-train --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], d['clocks'])"
done
