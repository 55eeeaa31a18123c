#!/bin/bash
# kernel probes + micro-benchmarks + GPU parity tests + bench line
mkdir -p gpurun_out
echo "== probes"; timeout 600 python tests/gpu_probe.py "$@" 2>&1 | tail -60 | tee gpurun_out/probe.log
echo "== elem bench"; timeout 300 python tests/elem_bench.py 2>&1 | tail -20 | tee gpurun_out/elem_bench.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
