#!/bin/bash
# two epilogue warp sets: parity + A/B timings
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_n.log
{
for e in 1 2; do
echo "== conv default EPI_SETS=$e"; VINCE_B200_EPI_SETS=$e python tests/conv_bench.py --filter "r" --iters 7
echo "== conv tstats EPI_SETS=$e";  VINCE_B200_EPI_SETS=$e python tests/conv_bench.py --filter "r50.layer" --iters 7 --tstats
echo "== conv apply 0 EPI_SETS=$e"; VINCE_B200_EPI_SETS=$e python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 0
echo "== conv apply 1 EPI_SETS=$e"; VINCE_B200_EPI_SETS=$e python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
done
} 2>&1 | tee gpurun_out/conv_variants_n.log
for v in "1 0" "2 0" "2 1"; do set -- $v
for c in 2 1; do
echo "== bench --config $c EPI_SETS=$1 TWOPASS=$2"; VINCE_B200_EPI_SETS=$1 VINCE_B200_TWOPASS=$2 timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-train --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'nce', d['infonce_step_ms'], d['infonce_step_with_dq_backward_ms'], d['clocks'])"
done
done
