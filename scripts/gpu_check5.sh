#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== nce bench"; timeout 300 python tests/elem_bench.py --only nce 2>&1 | tail -5 | tee gpurun_out/nce_bench.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
