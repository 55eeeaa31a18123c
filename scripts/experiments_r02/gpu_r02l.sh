#!/bin/bash
# transposed statistics pass + apply epilogue: parity, per-layer timings, step A/B
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -s -x -k "transposed or fused_epilogue or batch_separable or knn" 2>&1 | grep -E "rel|passed|failed|rror|assert|FAILED|stats_only" | tail -30 | tee gpurun_out/pytest_gpu_l.log
{
echo "== conv default";   python tests/conv_bench.py --filter "r50.layer" --iters 7
echo "== conv tstats";    python tests/conv_bench.py --filter "r50.layer" --iters 7 --tstats
echo "== conv tstats single"; VINCE_B200_TSTATS_PAIR=0 python tests/conv_bench.py --filter "r50.layer" --iters 7 --tstats
echo "== conv apply 0";   python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 0
echo "== conv apply 1";   python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 1
echo "== conv apply 2";   python tests/conv_bench.py --filter "r50.layer" --iters 7 --apply 2
} 2>&1 | tee gpurun_out/conv_variants_l.log
for v in "0 0" "1 0" "0 1" "1 1"; do set -- $v
echo "== bench --config 2 GRAPH=$1 TWOPASS=$2"; VINCE_B200_GRAPH=$1 VINCE_B200_TWOPASS=$2 timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
for g in 0 1; do
echo "== bench --config 1 GRAPH=$g"; VINCE_B200_GRAPH=$g timeout 600 python bench.py --config 1 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
echo "== pytest rest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
