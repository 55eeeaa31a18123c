#!/bin/bash
mkdir -p gpurun_out
echo "== probes"; timeout 600 python tests/gpu_probe.py 2>&1 | grep -v PASS | tail -20 | tee gpurun_out/probe.log
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -s -k "fused_epilogue or full_size or golden or cfg4 or fp64 or smoke" 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|cfg4|assert|FAILED" | tail -30 | tee gpurun_out/pytest_gpu_d.log
{
echo "== conv default"; python tests/conv_bench.py --filter "r50." --iters 7
echo "== conv staging=1"; VINCE_B200_STAGING=1 python tests/conv_bench.py --filter "r50." --iters 7
echo "== conv statsonly"; python tests/conv_bench.py --filter "r50." --iters 7 --statsonly
echo "== r18"; python tests/conv_bench.py --filter "r18." --iters 7
} 2>&1 | tee gpurun_out/conv_variants_d.log
for tp in 0 1; do
echo "== bench --config 2 TWOPASS=$tp"; VINCE_B200_TWOPASS=$tp timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
echo "== bench --config 1"; timeout 600 python bench.py --config 1 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
