#!/bin/bash
# One gpurun call: kernel probes, micro-benchmarks, GPU parity tests, smoke and the bench line.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_check.sh [probe filters...]'
mkdir -p gpurun_out
echo "== probes"; timeout 600 python tests/gpu_probe.py "$@" 2>&1 | grep -v PASS | tail -20 | tee gpurun_out/probe.log
echo "== conv bench"; timeout 300 python tests/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench.log
echo "== elem bench"; timeout 300 python tests/elem_bench.py 2>&1 | tail -24 | tee gpurun_out/elem_bench.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "rel-L2|passed|failed|rror" | tail -20 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
