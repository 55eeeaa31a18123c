#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x -k "fused_epilogue or full_size or golden" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_q.log
{
echo "== conv apply 1"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
echo "== conv apply 2"; python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 2
} 2>&1 | tee gpurun_out/conv_variants_q2.log
for v in "1"; do
echo "== bench --config 2 TWOPASS=$v"; VINCE_B200_TWOPASS=$v timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --no-train --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], d['clocks'])"
done
